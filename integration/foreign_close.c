/* Registers the host application's own pbf_close (the reference's, renamed at compile time) for handles that
 * were created by its writer (pbf_open_w), see INTEGRATION.md. */
#include "../include/pbwt_b200.h"
int ref_pbf_close(pbf_t *pb);
__attribute__((constructor)) static void b200_register_foreign_close(void) { pbf_b200_set_foreign_close(ref_pbf_close); }
