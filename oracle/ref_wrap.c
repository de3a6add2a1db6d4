/* ref_wrap.c — compiles the UNMODIFIED reference translation unit in place
 * (#include of /root/reference/krep.c, located with -I by build_oracle.py; no
 * reference source is copied into this repository) and adds three accessors
 * for the file-static option globals the kernels read (krep.c:117-120), which
 * no other translation unit can reach.
 *
 * TEST INFRASTRUCTURE ONLY — see oracle/krep_oracle.c header.
 */
#include "krep.c" /* resolved through -I$KREP_REF_DIR */

void krep_ref_set_only_matching(bool on) { only_matching = on; }
bool krep_ref_get_only_matching(void) { return only_matching; }
void krep_ref_set_force_no_simd(bool on) { force_no_simd = on; }
void krep_ref_set_algo_override(const char *name) { algo_override = name; }
