/* oracle/adapter_shim.h -- TEST INFRASTRUCTURE ONLY.
 * Force-included in front of the reference translation units of the ADAPTER build (oracle/make_adapter.py): ref_shim.h plus a
 * pinned time(), because the adapter keeps Terrain.cpp:58 (`static unsigned int randomseed = (unsigned int)time(NULL)`) and
 * hands that seed to rn_terrain2d. */
#ifndef RN_ADAPTER_SHIM_H
#define RN_ADAPTER_SHIM_H
#include "ref_shim.h"
#ifdef __cplusplus
extern "C" {
#endif
time_t rn_adapter_time(time_t*);
#ifdef __cplusplus
}
#endif
#define time rn_adapter_time
#endif
