/* Plain-C (C99) consumer of include/docscan.h: proves the boundary is a C ABI — the header compiles without C++ and the
 * library links and runs from C.  Built and run by tests/test_capi_cpu.py (no GPU needed: docscan_create must fail loudly). */
#include <stdio.h>
#include "docscan.h"
int main(void) {
    docscan_ctx* ctx = NULL;
    int rc = docscan_create(0, NULL, &ctx);
    printf("version %d create rc %d (%s)\n", docscan_version(), rc, docscan_strerror(rc));
    float quad[8] = {3.5f, 2.25f, 38.f, 4.f, 36.5f, 33.f, 1.f, 35.5f}, dst[8] = {0, 0, 29, 0, 29, 39, 0, 39};
    double m[9];
    rc = docscan_get_perspective_transform(quad, dst, m);
    printf("M[0]=%.17g rc %d\n", m[0], rc);
    docscan_params p; docscan_default_params(&p);
    printf("block %d canny %.0f/%.0f\n", p.block_size, p.canny_low, p.canny_high);
    int32_t per[180] = {0}; per[85] = 3; per[86] = 1; double a;
    docscan_median_angle(per, 10.0, &a);
    printf("angle %.17g\n", a);
    return 0;
}
