// Empty stand-in (TEST INFRASTRUCTURE): the compiled subset of the reference (vertex / half-plane contact, see
// oracle/ref_sym.cpp) uses nothing from the header of this name; the real one needs Eigen + muda, absent from this image.
#pragma once
