/* Minimal C caller of the batched entry (build: gcc examples/dff_batch_example.c -Iinclude
 * -Lraytracerfortran_b200 -lraytrace_b200 -Wl,-rpath,$PWD/raytracerfortran_b200 -o dff_batch_example).
 * Two 2-interface models, three sources; prints travel times and the fused logL.  Needs a B200. */
#include <stdio.h>
#include "raytrace_b200.h"

int main(void) {
    const double vels[2][3]   = {{2000.0, 3000.0, 4500.0}, {2500.0, 2600.0, 5000.0}};
    const double depths[2][2] = {{800.0, 2000.0}, {1200.0, 2500.0}};
    const int    nlayers[2]   = {2, 2};
    const double off[3] = {500.0, 1500.0, 3000.0}, dep[3] = {1000.0, 1800.0, 2600.0};
    const double tobs[3] = {0.5, 0.8, 1.1}, sigma[2] = {0.02, 0.03};
    double timeP[2][3], logL[2];
    const int B = 2, ldv = 3, ldz = 2, nsrc = 3;
    int rc = dff_batch(&vels[0][0], &depths[0][0], nlayers, &B, &ldv, &ldz, off, dep, &nsrc,
                       &timeP[0][0], tobs, sigma, logL, NULL);
    if (rc) {
        fprintf(stderr, "dff_batch failed: %s\n", rtb200_last_error());
        return 1;
    }
    for (int b = 0; b < B; ++b)
        printf("model %d: T = %.9f %.9f %.9f  logL = %.6f\n", b, timeP[b][0], timeP[b][1], timeP[b][2], logL[b]);
    rtb200_shutdown();
    return 0;
}
