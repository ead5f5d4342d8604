/* Stage B (VLIDORT) part of the C ABI -- included by aeroapps_b200.h */
#ifndef AEROAPPS_B200_RT_H
#define AEROAPPS_B200_RT_H
#endif
