/* TEST INFRASTRUCTURE ONLY. Compiles the reference's own lib/*.c (included from -I/root/reference, never
 * copied) into a shared object so tests can call fe_modp_mul, ec_jacobi_mulrdc, addr33, blf_has ... by their
 * reference names through ctypes.  main.c is pulled in for batch_add/check_found_add; its main() is renamed. */
#define main ecloop_ref_main
#include "main.c"
#undef main
