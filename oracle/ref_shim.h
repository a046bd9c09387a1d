/* oracle/ref_shim.h -- TEST INFRASTRUCTURE ONLY (see oracle/Makefile and DESIGN.md section 2).
 *
 * Force-included (g++ -include) in front of every UNMODIFIED reference
 * translation unit so that the Win32/MSVC-2008 sources under
 * /root/reference/renderEngine compile on Linux and draw their entropy from the
 * pinned MSVC CRT generator instead of glibc's.  Nothing here changes the
 * arithmetic the reference performs.
 *
 *  - `float random()` (reference include/basics.h:16) collides with glibc's
 *    `long random(void)`: include the libc headers FIRST, then rename.
 *  - RAND_MAX must be MSVC's 0x7fff or random() (basics.h:18) is ~-1 always.
 *  - rand/srand are renamed (not interposed) so the binding does not depend on
 *    the dynamic linker's symbol order when the .so is dlopen()ed from python.
 */
#ifndef RN_REF_SHIM_H
#define RN_REF_SHIM_H
#include <stdlib.h>
#include <stdio.h>
#include <math.h>
#include <time.h>
#include <string.h>
#ifdef __cplusplus
#include <new>
#include <cstdlib>
#include <cmath>
extern "C" {
#endif
int  rn_msvc_rand(void);
void rn_msvc_srand(unsigned int);
#ifdef __cplusplus
}
#endif
#define random rn_ref_random
#define rand   rn_msvc_rand
#define srand  rn_msvc_srand
#undef  RAND_MAX
#define RAND_MAX 0x7fff
#endif
