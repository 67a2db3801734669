/* ref_sheng_layout.c -- sizeof/offsetof of the Sheng structures, printed into
 * ref_layout_dump()'s JSON (ref_driver.c).  A separate translation unit because
 * src/nfa/sheng_internal.h and src/nfa/mcclellan_internal.h both define
 * struct report_list.  TEST INFRASTRUCTURE ONLY. */
#include <stddef.h>
#include <stdio.h>

#include "ue2common.h"
#include "nfa/sheng_internal.h"

#define SZ(s) printf("  \"sizeof(%s)\": %zu,\n", #s, sizeof(struct s))
#define OFF(s, f) printf("  \"%s.%s\": %zu,\n", #s, #f, offsetof(struct s, f))

void ref_layout_dump_sheng(void) {
    SZ(sheng);
    OFF(sheng, shuffle_masks); OFF(sheng, length); OFF(sheng, aux_offset);
    OFF(sheng, report_offset); OFF(sheng, accel_offset); OFF(sheng, n_states);
    OFF(sheng, anchored); OFF(sheng, floating); OFF(sheng, flags); OFF(sheng, report);
    SZ(sstate_aux);
    OFF(sstate_aux, accept); OFF(sstate_aux, accept_eod); OFF(sstate_aux, accel); OFF(sstate_aux, top);
}
