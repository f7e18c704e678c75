// bsw_internal.h -- entry points shared between the translation units of libbsw.so (not part of the C ABI).
#pragma once
#include "../../include/bsw.h"

extern "C" {
// max_ins/max_del as they arrive in the task batch buffer (param words 5,6: sw_pe_array_proc_element.v:924-934)
typedef struct { int32_t max_ins[2], max_del[2]; } bsw_seed_clamp;
// level 2 with optional host-supplied band clamps (clamps == NULL: ksw_extend2's formula)
int bsw_chain2aln_impl(bsw_ctx* ctx, const bsw_params2* P, const bsw_seed_task* tasks, size_t n,
                       const bsw_seed_clamp* clamps, bsw_aln_record* out);
void bsw_set_error_text(bsw_ctx* ctx, const char* text);
// value of an integer option the wire layer looks at ("fpga_strict"); -1 = unknown
int bsw_option_value(bsw_ctx* ctx, const char* key);
}
