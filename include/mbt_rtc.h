/*
 * mbt_rtc.h -- what the shared headers need from <stdint.h> / <math.h> / <string.h>, for the two ways they are compiled:
 * by nvcc / gcc (host headers available), and by NVRTC at run time (libmbt_b200 specialises its kernels for the handle's
 * configuration, see csrc/mbt_jit.h), where no host header exists.
 */
#ifndef MBT_RTC_H
#define MBT_RTC_H

#if defined(__CUDACC_RTC__)
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef short int16_t;
typedef unsigned short uint16_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
typedef long long intptr_t;
typedef unsigned long long uintptr_t;
#ifndef INFINITY
#define INFINITY __int_as_float(0x7f800000)
#endif
#ifndef NAN
#define NAN __int_as_float(0x7fffffff)
#endif
#else
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#endif

#endif /* MBT_RTC_H */
