/* Latency of dff_ (one model, 20 sources: BASELINE.json configs[0]) called from C, the way R's
 * .Fortran and a Fortran CALL reach it: no Python in the loop.  Next to it the CPU restatement
 * (oracle/liboracle_raymod.so, orc_trace_rays) for the same call.
 *   gcc -O2 profiles/latency_c.c -o /tmp/latency_c -Lraytracerfortran_b200 -lraytrace_b200 \
 *       -Loracle -loracle_raymod -Wl,-rpath,$PWD/raytracerfortran_b200 -Wl,-rpath,$PWD/oracle -lm */
#include <stdio.h>
#include <string.h>
#include <time.h>
#include "../include/raytrace_b200.h"

void orc_trace_rays(const double *v, const double *z, int nl, const double *so, const double *sd, int ns,
                    double *t, double *p, void *tr, int keep_delta, const char *path);

static double now_us(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

int main(void) {
    /* test_1: test_1_map.dat model, test_1_src_data.txt sources (15 digits, as shipped) */
    const double v[7] = {3100, 2400, 3730, 4200, 5800, 4420, 6200};
    const double z[6] = {1000, 1500, 2000, 2500, 3260, 3960};
    const double so[20] = {3011.895663618188, 5653.904846082881, 4446.623430480743, 4703.217011264789, 4058.4736793181496, 1802.6439402265723, 3709.519621973642, 1252.695362722442, 2945.1428258202377, 5581.584265690676, 2376.9317331436555, 6032.5723750762345, 6045.641472015896, 5489.823210673838, 4813.480930499342, 7518.724909112918, 5167.009737165377, 6065.4258728966, 3825.159979354237, 4500.570223675418};
    const double sd[20] = {1268.4868847834878, 3625.9918759809807, 4019.2584569449537, 1898.5529104596935, 1583.4465882391669, 1156.7712107906118, 1613.1727631902322, 3071.2459014728665, 1122.064891923219, 1076.223203795962, 2286.9961710297503, 3613.723761762958, 2235.1826344267465, 2249.5583787211217, 1884.4928895123303, 2433.903096697759, 2491.462526412215, 2753.2287737354636, 3146.8914641649462, 1405.0016532302834};
    double t[20], tc[20], pc[20];
    const int NL = 6, NS = 20, keep = -1;
    for (int i = 0; i < 200; ++i) dff_(v, z, &NL, so, sd, &NS, t, &keep);
    const int n = 20000;
    double t0 = now_us();
    for (int i = 0; i < n; ++i) dff_(v, z, &NL, so, sd, &NS, t, &keep);
    const double gpu = (now_us() - t0) / n;
    char trace[20 * 128];
    t0 = now_us();
    for (int i = 0; i < n; ++i) orc_trace_rays(v, z, NL, so, sd, NS, tc, pc, trace, -1, NULL);
    const double cpu = (now_us() - t0) / n;
    t0 = now_us();
    for (int i = 0; i < n; ++i) orc_trace_rays(v, z, NL, so, sd, NS, tc, pc, trace, -1, "/tmp/rays_latency.dat");
    const double cpuf = (now_us() - t0) / n;
    int same = memcmp(t, tc, sizeof t) == 0;
    printf("{\"workload\": \"config1 shape: dff_ on the test_1 model (6 interfaces) x 20 sources, one call, from C\", "
           "\"gpu_us_per_call\": %.2f, \"cpu_port_us_per_call\": %.2f, "
           "\"cpu_port_with_rays_dat_truncate_us_per_call\": %.2f, \"bit_identical\": %s, \"error\": \"%s\"}\n",
           gpu, cpu, cpuf, same ? "true" : "false", rtb200_last_error());
    return same ? 0 : 1;
}
