// Empty stand-in (TEST INFRASTRUCTURE): ccd.inl includes it, the instantiated point_triangle_ccd uses nothing from it.
#pragma once
