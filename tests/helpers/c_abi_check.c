/* The boundary must be a plain C ABI: this file is compiled as C99 with -Wall -Wextra -pedantic -Werror
 * by tests/test_capi_symbols.py. It only checks that the header is valid C and that the struct
 * layouts are what the reference-side binding assumes. */
#include <stddef.h>

#include "../../include/tuber_b200.h"

typedef char check_transform[(sizeof(tb_transform) == 48) ? 1 : -1];
typedef char check_sprite[(sizeof(tb_sprite) == 16) ? 1 : -1];
typedef char check_velocity[(sizeof(tb_velocity) == 12) ? 1 : -1];
typedef char check_instance[(sizeof(tb_instance) == 80) ? 1 : -1];
typedef char check_vertex[(sizeof(tb_quad_vertex) == 16) ? 1 : -1];
typedef char check_batch[(sizeof(tb_batch) == 16) ? 1 : -1];
typedef char check_frame_out[(offsetof(tb_frame_out, path) == 48 && sizeof(tb_frame_out) == 72) ? 1 : -1];

int c_abi_check(void) {
    tb_ctx* ctx = NULL;
    tb_frame_out out;
    /* never executed with a context: only has to compile and link against the declarations */
    if (ctx != NULL) {
        (void)tb_frame_transform_batch(ctx, &out);
        (void)tb_system_integrate(ctx, 0.01f);
    }
    return (int)sizeof(out);
}
