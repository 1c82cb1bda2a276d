// Stand-in for muda/muda_def.h (TEST INFRASTRUCTURE): the function qualifiers the host build of the reference headers needs.
#pragma once
#define MUDA_INLINE inline
#define MUDA_GENERIC
#define MUDA_HOST
#define MUDA_DEVICE
