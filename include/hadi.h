/* hadi.h — C ABI of the B200-native batched Heston ADI pricing / calibration engine.
 *
 * Drop-in boundary for the solver and calibration entry points of
 * BCW-dot/PDE-based-Heston-Solver-GPU-accelerated (citations relative to that repository):
 *
 *   reference entry point (host function that launches one Kokkos kernel)       replaced by
 *   ---------------------------------------------------------------------------------------------
 *   compute_base_prices{,_american,_dividends,_american_dividends}              hadi_price_batch
 *       src/jacobian_computation.hpp:82-104,127-143,168-189,216-237
 *   compute_base_prices_multi_maturity{,_american_dividends}                    hadi_price_batch
 *       src/heston_calibration.cpp:2339,3140                                    (per-point N, dt)
 *   compute_jacobian{,_american,_dividends,_american_dividends}                 hadi_jacobian_batch
 *       src/jacobian_computation.hpp:43-80,106-125,145-166,191-214
 *   compute_jacobian_multi_maturity{,_american_dividends}                       hadi_jacobian_batch
 *       src/heston_calibration.cpp:2174,2936
 *   compute_parameter_update_on_device  src/jacobian_computation.hpp:36-41      hadi_lm_update
 *   solve_5x5_device                    src/jacobian_computation.hpp:28-33      hadi_solve5
 *   LM loops of test_calibration_*      src/heston_calibration.cpp:204-417,     hadi_calibrate
 *                                       2692-2831, 3568-3716
 *   CalibrationPoint                    src/heston_calibration.cpp:2165-2171    hadi_point
 *   parallel_DO_solve                   src/device_solver.hpp:53                hadi_price_batch
 *   Jacobian with the V0 column interpolated on the base solve (prototype)      hadi_jacobian_batch_ex,
 *       src/device_solver.cpp:1725-1829                                         hadi_calibrate_ex (opt-in)
 *   BlackScholes::call_vega / reverse_BS / reverse_BS_dic  src/bs.hpp:124-192   hadi_bs_vega, hadi_bs_implied_vol{,_bisect}
 *   BlackScholes::generate_market_data{,_with_dividends}   src/bs.hpp:58-112    hadi_market_prices, hadi_dividend_adjusted_spot
 *   implied-vol post-processing and CSV export of the LM drivers                hadi_implied_vols, hadi_write_calibration_csv
 *       src/heston_calibration.cpp:436-511, 2853-2923
 *   CS_scheme_shuffled / MCS_scheme_shuffled (host solver)                      hadi_price_batch with hadi_numerics::scheme
 *       src/solver.hpp:781-907, 917-1075                                        (large grids: the wide kernel)
 *   ConvergenceExporter::testWithRelatedGridSizes / exportToCSV                 hadi_convergence_study,
 *       src/solver.cpp:50-295                                                   hadi_write_convergence_csv
 *   (no counterpart) the LM loop on one solver call per iteration               hadi_calibrate_ex with
 *                                                                               HADI_LM_SCHEDULE_SPECULATIVE (opt-in)
 *
 * Conventions
 *   - Plain C types only.  All pointers are HOST pointers unless the name ends in _dev.
 *   - Grids are built inside the library exactly as every reference caller builds them:
 *     Grid(m1, 8K, S0, K, K/5, m2, 5.0, V0, 5.0/500)  (src/grid.cpp:16, src/heston_calibration.cpp:2614),
 *     the variance grid being rebuilt for the current V0 as GridViews::rebuild_variance_views does
 *     (src/grid_pod.hpp:25).  The payoff U_0 the reference takes as an input array is generated from
 *     `payoff` (max(s-K,0) or max(K-s,0)); boundary vectors are the reference's call vectors.
 *   - Jacobian / delta column order: (kappa, eta, sigma, rho, v0)  (src/jacobian_computation.cpp:299-304,332).
 *   - Every function returns HADI_OK (0) or a negative error code; nothing throws across the ABI.
 *   - A hadi_ctx is bound to one CUDA device and must be used from one host thread at a time.
 *   - There is NO CPU fallback: if no CUDA device is usable hadi_create fails with HADI_ERR_CUDA.
 */
#ifndef HADI_H
#define HADI_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HADI_OK 0
#define HADI_ERR_ARG (-1)        /* bad shape / null pointer / unsupported combination */
#define HADI_ERR_GRID (-2)       /* S0 is not a node of some option's s-grid (reference: index_s = -1, UB) */
#define HADI_ERR_CUDA (-3)       /* CUDA runtime error (see hadi_last_error) */
#define HADI_ERR_SMEM (-4)       /* grid does not fit the shared-memory resident kernel */
#define HADI_ERR_COMM (-5)       /* multi-GPU exchange failed */
#define HADI_ERR_NOMEM (-6)

#define HADI_EUROPEAN 0
#define HADI_AMERICAN 1
#define HADI_CALL 0
#define HADI_PUT 1               /* put PAYOFF under the reference's call boundary vectors */
#define HADI_DOUGLAS 0
#define HADI_CRAIG_SNEYD 1
#define HADI_MODIFIED_CRAIG_SNEYD 2   /* MCS_scheme_shuffled AS SHIPPED (src/solver.hpp:917-1075; the reference marks it as not
                                         working and it is not a usable pricer: 3.76 where the other schemes give 8.87) —
                                         reproduced bit for bit, for callers that depend on the reference's numbers */
#define HADI_HUNDSDORFER_VERWER 3     /* extension, not in the reference (parity unpinned): second order in time for any theta */
/* opt-in extensions beyond the reference's device path.  PARITY UNPINNED: the reference
 * has no such code path; the CPU restatement in oracle/hadi_oracle.c is the definition the GPU path is tested against. */
#define HADI_BC_REFERENCE_CALL 0   /* the reference's boundary vectors b1, b2 (call asymptotics; src/BoundaryConditions.hpp:7-12) */
#define HADI_BC_PUT 1              /* put-correct set: b1 = b2 = 0 and Dirichlet U(s_0, v, tau) = K exp(-r_d tau) after every step */
#define HADI_DIVIDENDS_DEVICE 0    /* the device path's schedule: one dividend per step at most (quirk Q7, src/device_solver.hpp:432-516) */
#define HADI_DIVIDENDS_ALL 1       /* the host solver's schedule: every dividend dated inside the step, in order (src/solver.hpp:363) */

typedef struct hadi_ctx hadi_ctx;
typedef struct hadi_batch hadi_batch;

/* Market + model parameters: the scalar arguments S_0, V_0, r_d, r_f, rho, sigma, kappa, eta of
 * compute_base_prices / compute_jacobian (src/jacobian_computation.hpp:45-47). */
typedef struct {
  double S0, V0, r_d, r_f;
  double kappa, eta, sigma, rho;
} hadi_model;

/* One option to price.  Same fields and meaning as the reference's CalibrationPoint
 * (src/heston_calibration.cpp:2165-2171); results are written at position global_index. */
typedef struct {
  double strike;
  double maturity;
  int time_steps;   /* N */
  double delta_t;   /* maturity / N as computed by the caller */
  int global_index;
} hadi_point;

/* Numerical set-up shared by a batch: m1, m2, theta of the reference signatures, the variant
 * (which of the four reference functions), and the dividend schedule (dates in time-to-maturity
 * units as the reference compares them with t = n*dt: src/device_solver.hpp:432-516). */
typedef struct {
  int m1, m2;
  double theta;
  int style;        /* HADI_EUROPEAN | HADI_AMERICAN */
  int payoff;       /* HADI_CALL | HADI_PUT */
  int scheme;       /* HADI_DOUGLAS (device path of the reference) | HADI_CRAIG_SNEYD | HADI_MODIFIED_CRAIG_SNEYD |
                       HADI_HUNDSDORFER_VERWER (the last three: European, no dividends, reference call boundary vectors) */
  int num_dividends;
  const double* dividend_dates;
  const double* dividend_amounts;
  const double* dividend_percentages;
  int boundary;            /* HADI_BC_REFERENCE_CALL (0, the parity path) | HADI_BC_PUT (Douglas only) */
  int dividend_schedule;   /* HADI_DIVIDENDS_DEVICE (0, the parity path) | HADI_DIVIDENDS_ALL */
} hadi_numerics;

typedef struct {
  int max_iter;      /* reference: 15 (multi-maturity), 20 (American dividends) */
  double tol;        /* stop when sum r^2 < tol          (src/heston_calibration.cpp:2544) */
  double delta_tol;  /* stop when ||delta||_2 < delta_tol (:2545) */
  double lambda0;    /* 0.01 (:2676) */
  double eps;        /* finite-difference bump, 1e-6 (:2457) */
} hadi_lm_options;

typedef struct {
  double params[5];  /* kappa, eta, sigma, rho, v0 */
  double final_error;
  double lambda;
  double delta_norm;
  int iterations;
  int converged;
  int pde_solves;
  double gpu_ms;     /* device time of all solver launches (CUDA events) */
  long long exact_reruns;  /* solves repeated with IEEE divisions (see hadi_exact_reruns); 0 on option data so far */
} hadi_lm_result;

/* ---- context ------------------------------------------------------------------------------- */
int hadi_create(hadi_ctx** out, int device);
void hadi_destroy(hadi_ctx* ctx);
const char* hadi_last_error(const hadi_ctx* ctx);
const char* hadi_version(void);
/* number of hadi kernels launched by this context so far (bench.py's gpu_launches) */
long long hadi_kernel_launches(const hadi_ctx* ctx);
/* Solves this context had to repeat with IEEE divisions because a guarded fast division left its operand range
 *  or a split-schedule hand-off timed out.  The published values are exact either way; a
 * non-zero count means those solves cost twice the time.  Updated by every hadi_batch_fetch / one-call entry point. */
long long hadi_exact_reruns(const hadi_ctx* ctx);
/* the same for the last fetched launch of one batch */
long long hadi_batch_exact_reruns(const hadi_batch* b);
/* cumulative host->device / device->host bytes moved by this context */
int hadi_transfer_bytes(const hadi_ctx* ctx, long long* h2d, long long* d2h);

/* ---- one-call entry points (host buffers in, host buffers out; synchronous) ------------------- */
/* prices[n] are written at points[k].global_index (as the reference's multi-maturity drivers do); the optional
 * U_out[n*(m1+1)*(m2+1)] and lambda_out (same shape; may be NULL) are in INPUT order: block k belongs to points[k]. */
int hadi_price_batch(hadi_ctx* ctx, const hadi_model* model, const hadi_numerics* num, int n,
                     const hadi_point* points, double* prices, double* U_out, double* lambda_out);
/* J[n*5] row-major, base_prices[n]. */
int hadi_jacobian_batch(hadi_ctx* ctx, const hadi_model* model, const hadi_numerics* num, int n,
                        const hadi_point* points, double eps, double* J, double* base_prices);

/* ---- prepared batches: descriptors + grids resident in HBM, launch / fetch separately --------- */
#define HADI_MODE_PRICE 0
#define HADI_MODE_JACOBIAN 1          /* the reference's Jacobian: forward differences, 6 solves per option */
/* opt-in alternatives; they change the numbers, so parity with the reference's
 * LM trajectory holds for HADI_MODE_JACOBIAN only:
 *   INTERP   the V0 column from the base solve, interpolated linearly in v between the rows bracketing
 *            V0 + eps (the reference's prototype, src/device_solver.cpp:1725-1829): 5 solves per option,
 *            item index = option*5 + {base, kappa, eta, sigma, rho}, THREE values per item
 *            {price, U(S0, v_lower), U(S0, v_upper)}
 *   CENTRAL  (p(+eps) - p(-eps)) / (2 eps): 11 solves per option, item index = option*11 +
 *            {base, +kappa, +eta, +sigma, +rho, +v0, -kappa, -eta, -sigma, -rho, -v0} */
#define HADI_MODE_JACOBIAN_INTERP 2
#define HADI_MODE_JACOBIAN_CENTRAL 3
/* How hadi_calibrate_ex spends its solver calls.  REFERENCE is the reference's loop (src/heston_calibration.cpp:
 * 2692-2831): Jacobian at the current point (6n solves), candidate prices (n solves), and after a rejected step the same
 * Jacobian again.  SPECULATIVE evaluates the candidate with the Jacobian batch itself — its base column IS the
 * candidate's prices — so that an accepted step already holds the next iteration's Jacobian, and keeps the Jacobian of
 * the current point across rejected steps: one solver call per iteration instead of two, 6n solves per iteration
 * instead of 7n, and the same parameters, errors and lambda bit for bit (every solve is the same solve; only
 * hadi_lm_result::pde_solves differs). */
#define HADI_LM_SCHEDULE_REFERENCE 0
#define HADI_LM_SCHEDULE_SPECULATIVE 1
typedef struct {
  int mode;        /* HADI_MODE_JACOBIAN | _INTERP | _CENTRAL */
  double eps[5];   /* bump per parameter (kappa, eta, sigma, rho, v0); the reference uses 1e-6 for all */
  int schedule;    /* HADI_LM_SCHEDULE_* (hadi_calibrate_ex only) */
} hadi_jacobian_options;
/* Work items are options (PRICE) or option x {base, kappa, eta, sigma, rho, v0} (JACOBIAN), item
 * index = option*6 + column.  [item_begin, item_end) selects the slice this context solves
 * (multi-GPU sharding); pass 0, -1 for everything. */
int hadi_batch_create(hadi_ctx* ctx, const hadi_model* model, const hadi_numerics* num, int n,
                      const hadi_point* points, int mode, double eps, int item_begin, int item_end,
                      hadi_batch** out);
/* as hadi_batch_create with one bump per parameter (eps5[5]) */
int hadi_batch_create_ex(hadi_ctx* ctx, const hadi_model* model, const hadi_numerics* num, int n,
                         const hadi_point* points, int mode, const double* eps5, int item_begin, int item_end,
                         hadi_batch** out);
int hadi_batch_num_items(const hadi_batch* b);
/* Which kernel the library planned for this batch (diagnostics, benchmarks): variant id (0 / 1
 * grid-specialised, 2 / 3 run-time dimensions, 5 / 6 one CTA per solve with the working set in global scratch, 7 one
 * thread-block cluster per solve, 9 the wide kernel: one solve on a team of co-resident CTAs), CTAs of the launch and
 * CTAs that share one solve.  Any pointer may be NULL. */
int hadi_batch_kernel_info(const hadi_batch* b, int* variant, int* grid_ctas, int* ctas_per_solve);
/* values each item publishes: 1, or 3 in HADI_MODE_JACOBIAN_INTERP */
int hadi_batch_values_per_item(const hadi_batch* b);
/* Re-aim a prepared batch at new Heston parameters (kappa, eta, sigma, rho, V0 of `model`; S0, r_d, r_f must be those
 * of creation) without rebuilding it: only the item descriptors and the v-grids are rewritten and uploaded.  The LM
 * driver uses it for the solver calls of one calibration; the previous launch must have been fetched. */
int hadi_batch_update_model(hadi_batch* b, const hadi_model* model);
/* enqueue the solve on the context's stream (asynchronous) */
int hadi_batch_launch(hadi_batch* b);
/* device pointer to the item values (one double per item of the slice, slice order) */
double* hadi_batch_values_dev(hadi_batch* b);
/* wait and copy the slice's item values to the host (values[(item_end-item_begin) * values_per_item]) */
int hadi_batch_fetch(hadi_batch* b, double* values);
/* elapsed device time of the last launch in ms (CUDA events on the context's stream) */
int hadi_batch_elapsed_ms(hadi_batch* b, float* ms);
/* development aid: per-phase SM cycle counters of the last launch (zeros unless built with
 * -DHADI_PHASE_TIMING); out8[8] */
int hadi_batch_phase_cycles(hadi_batch* b, long long* out8);
void hadi_batch_destroy(hadi_batch* b);

/* Jacobian assembly from the full [n*6] item values: base = v[6k], J[k][c] = (v[6k+1+c]-v[6k])/eps
 * (src/jacobian_computation.cpp:330,361). */
int hadi_jacobian_assemble(int n, const double* item_values, double eps, double* J, double* base_prices);

/* Jacobian rows from the item values of any Jacobian mode (layouts above); v0_weight from hadi_jacobian_v0_weight */
int hadi_jacobian_assemble_ex(int n, int mode, const double* item_values, const double* eps5, double v0_weight,
                              double* J, double* base_prices);
/* rows of the base v-grid bracketing V0 + eps_v0 and the linear weight (src/device_solver.cpp:1735-1754) */
int hadi_jacobian_v0_weight(int m2, double V0, double eps_v0, int* lower, int* upper, double* weight);
/* hadi_jacobian_batch with the Jacobian taken as `opt` says */
int hadi_jacobian_batch_ex(hadi_ctx* ctx, const hadi_model* model, const hadi_numerics* num, int n,
                           const hadi_point* points, const hadi_jacobian_options* opt, double* J,
                           double* base_prices);

/* Static block partition of item costs over `world` ranks (contiguous, balanced by cost). */
int hadi_partition(int n_items, const int* costs, int world, int rank, int* begin, int* end);
/* cost (N*P) of each item of a would-be batch, for hadi_partition; costs[n * items per option] */
int hadi_item_costs(const hadi_numerics* num, int n, const hadi_point* points, int mode, int* costs);

/* Inspection aid: the split schedule a batch of n solves of time_steps[k] steps gets on `slots` persistent CTAs.
 * When a batch holds more solves than CTAs, the one solve that straddles the end of a CTA's share is cut in two
 * (its first steps open the next CTA's list, its state is handed over through L2), so that all CTAs finish
 * together instead of after a whole number of solves.  seg5[5*q] = {item, first step, last step, hand-off in,
 * hand-off out}, slot_off[slots+1]; `setup` = cost of a segment's set-up in steps (the library uses 1.5).
 * Returns the number of segments, 0 when nothing is cut (n <= slots). */
int hadi_plan_schedule(int n, const int* time_steps, int slots, double setup, int max_segments, int* seg5,
                       int* slot_off, double* heaviest_steps);

/* ---- Levenberg-Marquardt ---------------------------------------------------------------------- */
int hadi_solve5(const double* A, const double* b, double* x);
int hadi_lm_update(int n, const double* J, const double* residual, double lambda, double* delta);

/* Optional exchange hook for multi-GPU calibration: called with this rank's slice of item values
 * (host buffer) and must fill `all` with every rank's slice in rank order (an all-gather).
 * counts/displs are in doubles.  Return 0 on success. */
typedef int (*hadi_allgather_fn)(void* user, const double* mine, int my_count, double* all,
                                 const int* counts, const int* displs, int world);
typedef struct {
  int rank, world;
  hadi_allgather_fn allgather;
  void* user;
} hadi_comm;

/* In-library exchange: an NCCL communicator owned by the context (one context = one device = one rank).  Rank 0
 * obtains an id with hadi_nccl_unique_id and hands the 128 bytes to every rank by any means (MPI, a file,
 * torch.distributed); every rank then calls hadi_comm_init.  With a communicator attached, hadi_calibrate /
 * hadi_calibrate_ex (comm == NULL) and the *_sharded entry points split the work items over the ranks
 * (hadi_partition), each rank's kernel publishes its values straight into its slot of a device gather buffer, and one
 * ncclAllGather on the context's stream plus one device-to-host copy per solver call return every value to every
 * rank — no host hop, no callback.  NCCL is loaded with dlopen("libnccl.so.2") on first use (HADI_NCCL_LIB overrides
 * the name); libhadi.so itself does not link it. */
#define HADI_NCCL_ID_BYTES 128
int hadi_nccl_unique_id(void* id128);
int hadi_comm_init(hadi_ctx* ctx, int world, int rank, const void* id128);
void hadi_comm_finalize(hadi_ctx* ctx);
int hadi_comm_world(const hadi_ctx* ctx);
int hadi_comm_rank(const hadi_ctx* ctx);
/* SPMD one-call entry points: every rank passes the same arguments and receives every result */
int hadi_price_batch_sharded(hadi_ctx* ctx, const hadi_model* model, const hadi_numerics* num, int n,
                             const hadi_point* points, double* prices);
int hadi_jacobian_batch_sharded(hadi_ctx* ctx, const hadi_model* model, const hadi_numerics* num, int n,
                                const hadi_point* points, const hadi_jacobian_options* opt, double* J,
                                double* base_prices);

/* Full LM calibration (host loop of src/heston_calibration.cpp:2692-2831 around the batched
 * solver).  comm may be NULL: single GPU, or the context's own communicator when one is attached (hadi_comm_init). */
int hadi_calibrate(hadi_ctx* ctx, const hadi_model* initial, const hadi_numerics* num, int n,
                   const hadi_point* points, const double* market_prices, const hadi_lm_options* opt,
                   const hadi_comm* comm, hadi_lm_result* result);

/* hadi_calibrate with the Jacobian taken as `jopt` says (opt->eps is ignored) */
int hadi_calibrate_ex(hadi_ctx* ctx, const hadi_model* initial, const hadi_numerics* num, int n,
                      const hadi_point* points, const double* market_prices, const hadi_lm_options* opt,
                      const hadi_jacobian_options* jopt, const hadi_comm* comm, hadi_lm_result* result);

/* ---- measurement ------------------------------------------------------------------------------ */
/* FP64 issue-rate micro-benchmark on `device` (CUDA events): un-fused DMUL+DADD rate and DFMA rate in
 * TFLOP/s, dependent-DADD latency in ns.  The un-fused rate is the roofline denominator of the
 * SMEM-resident kernel (parity forbids FMA).  Any output pointer may be NULL. */
int hadi_measure_fp64(int device, double* unfused_tflops, double* fma_tflops, double* dep_latency_ns);

/* ---- host helpers restating small reference utilities ---------------------------------------- */
/* Grid::Grid (src/grid.cpp:16-96) with the callers' constants; s[m1+1], v[m2+1] */
int hadi_grid(int m1, int m2, double K, double S0, double V0, double* s, double* v);
/* BlackScholes::call_price (src/bs.hpp:44-55) */
double hadi_bs_call(double S, double K, double r, double vol, double T);

/* ---- callers either side of the calibration driver (host only) ----- */
/* BlackScholes::call_vega, CP = 1 (src/bs.hpp:124-127) */
double hadi_bs_vega(double S, double K, double r, double vol, double T);
/* BlackScholes::reverse_BS_dic (src/bs.hpp:131-160): bisection on [a, b] until |C - target| <= epsilon */
double hadi_bs_implied_vol_bisect(double S, double K, double r, double T, double target, double epsilon,
                                  double a, double b);
/* BlackScholes::reverse_BS (src/bs.hpp:164-192): Newton from v0, bisection on [0.001, 1] when vega < 1e-10 */
double hadi_bs_implied_vol(double S, double K, double r, double T, double v0, double target, double epsilon);
/* spot net of the discounted dividends paid before T (src/bs.hpp:91-102, src/heston_calibration.cpp:1500-1511) */
double hadi_dividend_adjusted_spot(double S0, double T, double r_d, int num_dividends, const double* dates,
                                   const double* amounts, const double* percentages);
/* BlackScholes::generate_market_data{,_with_dividends} (src/bs.hpp:58-112) per calibration point;
 * prices[global_index]; num_dividends = 0 selects the plain generator */
int hadi_market_prices(double S0, double r_d, double market_vol, int n, const hadi_point* points,
                       int num_dividends, const double* dates, const double* amounts, const double* percentages,
                       double* prices);
/* implied vols of market and fitted prices and their absolute difference (src/heston_calibration.cpp:441-460,
 * 2890-2910); arrays indexed by global_index, any output may be NULL */
int hadi_implied_vols(double spot, double r_d, int n, const hadi_point* points, const double* market,
                      const double* fitted, double epsilon, double* market_iv, double* fitted_iv, double* iv_diff);
/* calibration report: format 0 = fitted_heston_vs_market.csv (src/heston_calibration.cpp:467-508),
 * format 1 = fitted_heston_vs_market_multi_maturity.csv (:2857-2921) */
int hadi_write_calibration_csv(const char* path, int format, double spot, double r_d, int num_maturities,
                               int num_strikes, const hadi_point* points, const double* market,
                               const double* fitted, const hadi_model* initial, const hadi_lm_result* result,
                               double total_time_s, double iv_epsilon);

/* ---- convergence-study harness (src/solver.cpp:50-295) ---------------------------- */
/* ConvergenceExporter::testWithRelatedGridSizes: one call (strike K, maturity T) priced on the grids m1 = 2*m2 for every
 * m2 of the list with N steps of `scheme`; relative error against ref_price, mean wall seconds of `repeats` solves
 * (the reference uses N = 20, theta = 0.8, 20 repeats).  Output arrays have n_sizes entries; any may be NULL. */
int hadi_convergence_study(hadi_ctx* ctx, const hadi_model* model, double K, double T, int N, double theta, int scheme,
                           int n_sizes, const int* m2_sizes, double ref_price, int repeats, double* prices,
                           double* rel_errors, double* seconds);
/* ConvergenceExporter::exportToCSV: header m1,m2,price,error,time; scientific notation, ten digits */
int hadi_write_convergence_csv(const char* path, int n_sizes, const int* m2_sizes, const double* prices,
                               const double* rel_errors, const double* seconds);

#ifdef __cplusplus
}
#endif
#endif /* HADI_H */
