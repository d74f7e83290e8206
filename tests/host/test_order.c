/* CPU check of nbh_order_by_id (necat_b200/host/nb_host.c): ascending ids, stable.  Built and run by tests/test_host.py. */
#include <stdio.h>
#include <stdlib.h>
#include "../../necat_b200/host/nb_host.h"

int main(void)
{
    const uint64_t sizes[] = { 0, 1, 2, 1000, 70001 };
    unsigned s = 12345;
    for (unsigned t = 0; t < sizeof sizes / sizeof sizes[0]; ++t) {
        const uint64_t n = sizes[t];
        nb_cns_record* r = (nb_cns_record*)calloc(n ? n : 1, sizeof *r);
        uint64_t* o = (uint64_t*)malloc(8 * (n ? n : 1));
        for (uint64_t i = 0; i < n; ++i) {
            s = s * 1103515245u + 12345u;
            r[i].id = (int32_t)((s >> 8) % (t == 3 ? 50u : 2000000000u));      /* many ties in the 1000 case */
            r[i].left = (int32_t)i;
        }
        if (nbh_order_by_id(r, n, o)) return 2;
        char* seen = (char*)calloc(n ? n : 1, 1);
        for (uint64_t k = 0; k < n; ++k) {
            if (o[k] >= n || seen[o[k]]) { printf("not a permutation\n"); return 1; }
            seen[o[k]] = 1;
            if (k && (r[o[k - 1]].id > r[o[k]].id || (r[o[k - 1]].id == r[o[k]].id && o[k - 1] > o[k]))) { printf("order broken at %lu\n", (unsigned long)k); return 1; }
        }
        free(r); free(o); free(seen);
    }
    printf("ok\n");
    return 0;
}
