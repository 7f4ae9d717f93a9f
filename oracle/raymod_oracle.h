/*
 * raymod_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the reference's 1-D layered-cake forward ray tracer
 * (module raymod, /root/reference/subroutineR-quiet.f90) and of the Gaussian
 * travel-time likelihood (LOGLHOOD_RT, /root/reference/ray_tracing_sampling/loglhood.f90).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (raytracerfortran_b200/) never does.
 *
 * Parity status: PINNED.  The restatement is checked against the reference's own
 * full-precision ray dump (rays.dat, 20 rays) and the rendered notebook's 5 printed
 * travel times (tests/test_oracle_golden.py).  The reference Fortran itself cannot be
 * compiled in this image (no Fortran compiler of any kind), so oracle/_ref/ is empty.
 *
 * Numerics contract: IEEE-754 binary64, round-to-nearest, NO fused multiply-add,
 * sums accumulated left to right from 0 -- what gfortran -O2 emits on baseline x86-64
 * for the array expressions of the reference.  Compile with -O2 -ffp-contract=off.
 */
#ifndef RAYMOD_ORACLE_H
#define RAYMOD_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* Which path GetPTime took for one ray (subroutineR-quiet.f90:94,139,142,145). */
enum {
    ORC_BRANCH_TOP    = 0,  /* source in the top layer: straight ray          */
    ORC_BRANCH_NEG    = 1,  /* f(p0) < 0           -> Newton from p0          */
    ORC_BRANCH_SAFE   = 2,  /* first Newton jump is safe -> Newton from p0    */
    ORC_BRANCH_BISECT = 3   /* bisection first, then Newton                   */
};

/* Per-ray trace of the solver trajectory.  Filled when a non-NULL pointer is given. */
typedef struct {
    double p;          /* ray parameter the travel time was summed at (p_final)          */
    double f_final;    /* last value of the offset misfit f seen by solve()              */
    int    nl;         /* 1-based index of the layer holding the source (whichLayer)     */
    int    branch;     /* ORC_BRANCH_*                                                   */
    int    n_halve;    /* p0 halvings in the NaN guard (:125-133)                        */
    int    n_bisect;   /* bisection iterations executed (loop trips of solvebst)         */
    int    n_newton;   /* Newton updates x <- x - f/f' applied in solve()                */
    int    n_clamp;    /* times solve() clamped x to 1/vmax - 1e-9 (:299-304)            */
    int    conv;       /* the reference's `conv` flag at exit (quirk at :327-330)        */
    int    n_f_ref;    /* costFunc calls the reference executes                          */
    int    n_fp_ref;   /* costFunc_Prime calls the reference executes                    */
    int    n_ffp_min;  /* distinct points where BOTH f and f' are consumed               */
    int    n_f_min;    /* distinct points where only f is consumed                       */
} orc_trace;

/* whichLayer, subroutineR-quiet.f90:9-32.  Returns the 1-based layer index. */
int orc_which_layer(const double *depths, int nlayers, double dph);

/* One ray: whichLayer + InsertLayer + GetPTime (subroutineR-quiet.f90:436-455).
 * vels[nlayers+1], depths[nlayers] (interface depths from the surface). */
double orc_ray_time(const double *vels, const double *depths, int nlayers,
                    double src_offset, double src_depth, orc_trace *tr);

/* dofullforwardproblem / TraceRays, subroutineR-quiet.f90:408-463 / :467-520.
 * p_out and traces may be NULL.  keep_delta > 0 rewrites ./rays_path like the
 * reference rewrites ./rays.dat (pass NULL for rays_path to skip the file). */
void orc_trace_rays(const double *vels, const double *depths, int nlayers,
                    const double *src_offset, const double *src_depth, int nsrc,
                    double *timeP, double *p_out, orc_trace *traces,
                    int keep_delta, const char *rays_path);

/* The Gaussian log-likelihood of LOGLHOOD_RT (loglhood.f90:165-166,193-203),
 * NMODE = 1, ICOV = 1, IAR = 0. */
double orc_loglhood_from_times(const double *tpred, const double *tobs, int ndat, double sigma);

/* Batched sweep over models (the workload shape of replica.f90 / configs 2-5).
 * vels[B][ldv], depths[B][ldz], nlayers[B]; sources shared by every model.
 * timeP[B][nsrc], p_out[B][nsrc], logL[B] may each be NULL; logL needs tobs, sigma[B].
 * nthreads <= 0 -> all OpenMP threads.  Returns the thread count used. */
int orc_dff_batch(const double *vels, const double *depths, const int *nlayers,
                  int B, int ldv, int ldz,
                  const double *src_offset, const double *src_depth, int nsrc,
                  double *timeP, const double *tobs, const double *sigma,
                  double *logL, double *p_out, int nthreads);

/* "Faithful" variant of the batched sweep: one model per call AND the per-call
 * open/truncate/close of rays.dat the reference performs (:432-433). */
int orc_dff_batch_faithful(const double *vels, const double *depths, const int *nlayers,
                           int B, int ldv, int ldz,
                           const double *src_offset, const double *src_depth, int nsrc,
                           double *timeP, const char *rays_path);

/* LOGLHOOD_RT's model mapping (loglhood.f90:127-146): k Voronoi nodes with P velocities
 * vp[0..k-1] and interface depths ziface[0..k-2]; k == 1 -> two equal velocities over one
 * fake interface at 9999.9.  Returns logL; tpred[ndat] (may be NULL) gets DpredRT. */
double orc_loglhood_rt(int k, const double *vp, const double *ziface,
                       const double *src_offset, const double *src_depth, int nsrc,
                       const double *tobs, double sigma, double *tpred);

/* "Next" row N1: INTERPLAYER_novar (loglhood.f90:214-295) -- sort the k Voronoi nodes by depth
 * with the reference's quicksort (quicksort.f90:66-123), in place. */
void orc_interplayer_novar(int k, double *node_depth, double *node_vp);
/* INTERPLAYER_novar followed by LOGLHOOD_RT on unsorted nodes; sorted_* (may be NULL) get the
 * sorted nodes. */
double orc_loglhood_voro(int k, const double *node_depth, const double *node_vp,
                         const double *src_offset, const double *src_depth, int nsrc,
                         const double *tobs, double sigma, double *tpred,
                         double *sorted_depth, double *sorted_vp);

/* IAR = 1: register the chains' AR(1) state (idxarRT, arparRT per chain, armxRT) so that the moves
 * below evaluate LOGLHOOD with the AR residual model (loglhood.f90:171-182); NULL = IAR = 0.  The
 * `chain` argument of the single-chain moves indexes these arrays (-1: no AR). */
void orc_set_chain_ar(const int *idxar, const double *arpar, double armx);

/* "Next" rows N1 + N2: one fixed-dimension MH move of one chain -- PROPOSAL (ENOS = 0 Cauchy step,
 * prjmh_temper_rf.f90:1386-1447), INTERPLAYER_novar, CHECKBOUNDS2 (:1681-1716), LOGLHOOD and the
 * accept test of EXPLORE_MH_NOVARPAR (:739-757).  Random numbers are inputs.  See the .c file. */
void orc_set_ismpprior(int on);   /* ISMPPRIOR = 1: every move's likelihood is LOGLHOOD2's constant 1 */
void orc_set_enos(int enos);   /* ENOS = 1: order-statistics prior in the move oracles (default 0) */
int orc_mh_step_chain(int chain, int k, double *node_depth, double *node_vp, double *logL,
                int ivo, int iwhich, double cauchy, double u_acc, double beta, double sigma,
                const double *prior,
                const double *src_offset, const double *src_depth, int nsrc, const double *tobs,
                double *prop_depth, double *prop_vp, double *logL_prop);
void orc_mh_step_batch(const int *k, double *voro, double *logL, int B, int ldk,
                       const int *ivo, const int *iwhich, const double *cauchy, const double *u_acc,
                       const double *beta, const double *sigma, const double *prior,
                       const double *src_offset, const double *src_depth, int nsrc,
                       const double *tobs, int *accept, double *voro_prop, double *logL_prop);

/* N2: the birth/death move of EXPLORE_MH_NOVARPAR (:658-710): move choice, BIRTH_FULL (:997-1103) or
 * DEATH_FULL (:917-994), CHECKBOUNDS (:1639-1678), LOGLHOOD, accept with the Poisson-prior logPr.
 * Returns 1 accepted, 0 rejected, -1 outside, 2 no birth/death proposed.  See the .c file. */
int orc_bd_step_chain(int chain, int *k_io, double *node_depth, double *node_vp, double *logL, int ldk,
                double u_k, int idel, double u_z, double u_v, double u_acc, double beta,
                double sigma, const double *prior, const double *pk, int kmin, int kmax,
                const double *src_offset, const double *src_depth, int nsrc, const double *tobs,
                int *k_prop, double *prop_depth, double *prop_vp, double *logL_prop);
void orc_bd_step_batch(int *k, double *voro, double *logL, int B, int ldk, const double *u_k,
                       const int *idel, const double *u_z, const double *u_v, const double *u_acc,
                       const double *beta, const double *sigma, const double *prior,
                       const double *pk, int kmin, int kmax,
                       const double *src_offset, const double *src_depth, int nsrc,
                       const double *tobs, int *accept, int *k_prop, double *voro_prop,
                       double *logL_prop);

/* N2: the data-error move of EXPLORE_MH (:545-575) with PROPOSAL_SDRT (:1616-1635). */
int orc_sd_step_chain(int chain, int k, const double *node_depth, const double *node_vp, double *logL, double *sigma,
                double u_gate, double gauss, double u_acc, double beta, const double *sd_prior,
                const double *src_offset, const double *src_depth, int nsrc, const double *tobs,
                double *logL_prop);
void orc_sd_step_batch(const int *k, const double *voro, double *logL, double *sigma, int B, int ldk,
                       const double *u_gate, const double *gauss, const double *u_acc,
                       const double *beta, const double *sd_prior,
                       const double *src_offset, const double *src_depth, int nsrc,
                       const double *tobs, int *accept, double *logL_prop);

/* N2 + N4: the AR(1) move of EXPLORE_MH (:583-631) with PROPOSAL_ARRT (:1521-1552). */
int orc_ar_step(int k, const double *node_depth, const double *node_vp, double *logL, double sigma,
                int *idxar, double *arpar, double u_choice, double u_prop, double gauss,
                double u_acc, double beta, const double *ar_prior,
                const double *src_offset, const double *src_depth, int nsrc, const double *tobs,
                double *logL_prop);
void orc_ar_step_batch(const int *k, const double *voro, double *logL, const double *sigma,
                       int *idxar, double *arpar, int B, int ldk, const double *u_choice,
                       const double *u_prop, const double *gauss, const double *u_acc,
                       const double *beta, const double *ar_prior,
                       const double *src_offset, const double *src_depth, int nsrc,
                       const double *tobs, int *accept, double *logL_prop);

/* "Next" row N4: LOGLHOOD_RT's likelihood with the AR(1) residual model of IAR = 1
 * (loglhood.f90:171-182, ARPRED_RT :616-653, CHECKBOUNDS_ARMXRT :678-699). */
double orc_loglhood_from_times_ar(const double *tpred, const double *tobs, int ndat, double sigma,
                                  int idxar, double arpar, double armx);

/* Aggregate trace counters over a batch (for W_ref / W_min flop accounting). */
typedef struct {
    long long rays, top, neg, safe, bisect;
    long long sum_nl, n_halve, n_bisect, n_newton, n_clamp, not_conv;
    long long n_f_ref, n_fp_ref, n_ffp_min, n_f_min;
    double    flops_ref, flops_min;   /* see DESIGN.md "work per evaluation" */
    long long sqrt_ref, div_ref, sqrt_min, div_min;
} orc_stats;

void orc_batch_stats(const double *vels, const double *depths, const int *nlayers,
                     int B, int ldv, int ldz,
                     const double *src_offset, const double *src_depth, int nsrc,
                     orc_stats *out);

#ifdef __cplusplus
}
#endif
#endif
