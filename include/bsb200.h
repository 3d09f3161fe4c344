/* bsb200.h -- C ABI of libbsb200.so: the B200-native replacement for the MCMC sampler hot path of
 * BayesSpace (reference: /root/reference, R package v1.15.3).
 *
 * Every entry point below replaces one symbol the reference registers for R's .Call interface
 * (src/RcppExports.cpp:166-175) or one helper those symbols call.  The reference-side binding (an
 * Rcpp shim that re-exports _BayesSpace_iterate* on top of this header) is in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; all buffers are caller-owned HOST memory unless stated;
 *   - matrices are column-major exactly as R stores them (Y is n x d, lambda0 is d x d);
 *   - labels are 1-based (as in R), neighbour ids are 0-based (as df_j is handed to C++,
 *     R/spatialCluster.R:289);
 *   - adjacency crosses the boundary as CSR: int64 ptr[n+1], int32 idx[nnz];
 *   - outputs are snapshot-major: snapshot r (r = 0 .. nrep/thin) is contiguous;
 *   - every function returns BSB200_OK (0) or a negative error code; bsb200_last_error() holds the
 *     message the reference would have raised as an R error (thread-local);
 *   - there is no CPU fallback: sampler and probe calls fail with BSB200_ERR_CUDA when no sm_100
 *     device is usable.
 */
#ifndef BSB200_H
#define BSB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSB200_VERSION 100 /* 0.1.0 */

enum {
  BSB200_OK = 0,
  BSB200_ERR_INVALID = -1,     /* "Invalid arguments!" (src/cluster.cpp:173) and argument checks  */
  BSB200_ERR_NEIGHBORS = -2,   /* "Error in finding neighbors of subspots!" (src/cluster.cpp:208)  */
  BSB200_ERR_PARSE = -3,       /* "Unconvertable value" / "Value out of range" (src/utils.h:39-41) */
  BSB200_ERR_BUFFER = -4,      /* output capacity too small; required size is reported             */
  BSB200_ERR_CUDA = -5,        /* CUDA runtime failure or no usable device                          */
  BSB200_ERR_NUMERIC = -6,     /* NA in a probability vector / Wishart df <= 0 / not pos. definite  */
                               /* (the reference errors inside sample() or chol())       */
  BSB200_ERR_UNSUPPORTED = -7, /* d > 32, q > 255, degree > 16, ...                                 */
  BSB200_ERR_CANCELLED = -8    /* bsb200_cancel() seen at a thin boundary (rows after it are zero)  */
};

enum { BSB200_MODEL_NORMAL = 0, BSB200_MODEL_T = 1 };
enum { BSB200_PRECISION_EQUAL = 0, BSB200_PRECISION_VARIABLE = 1 };
enum { BSB200_PLATFORM_VISIUM = 0, BSB200_PLATFORM_SQUARE = 1 /* "VisiumHD" and "ST" */ };
enum { BSB200_METRIC_MANHATTAN = 0, BSB200_METRIC_EUCLIDEAN = 1 };

/* Reference quirks reproduced by default. */
#define BSB200_COMPAT_Q1 1u /* covariance-form density uses (R R^T)^-1, src/cluster.cpp:69-99     */
#define BSB200_COMPAT_Q3 4u /* iterate_vvv compares neighbour indices to the label, :404          */
#define BSB200_COMPAT_REFERENCE (BSB200_COMPAT_Q1 | BSB200_COMPAT_Q3)

int bsb200_version(void);
const char *bsb200_last_error(void);
/* SIGTERM analogue (src/cluster.cpp:41-47, :1106-1107): running samplers stop at the next thin
 * boundary.  The flag is cleared when a sampler call starts. */
void bsb200_cancel(void);
/* User interrupts (Rcpp::checkUserInterrupt() every 10 iterations, src/cluster.cpp:243-244, :356-357, :475-476,
 * :606-607, :867-868): the host registers a poll function; every sampler call invokes it on the calling thread
 * after each batch of at most `thin` iterations, and a non-zero return stops the call like bsb200_cancel()
 * (BSB200_ERR_CANCELLED; rows_done tells which snapshot rows are complete).  The function must return, not
 * longjmp / throw (the R shim wraps R_CheckUserInterrupt in R_ToplevelExec).  NULL unregisters. */
typedef int (*bsb200_poll_fn)(void *user);
void bsb200_set_interrupt_poll(bsb200_poll_fn poll, void *user);
int bsb200_device_count(void);

/* Chain persistence (R/mcmcChain.R:48-73, :181-223): instead of materialising the snapshot matrices, a sampler call
 * hands every snapshot row to a registered sink as it is produced -- "z" (int32, n), "weights" (f64, n) and, for
 * bsb200_deconv, "Y" (f64, n x d flattened ROW-major like as.vector(t(Y))) at every thin boundary (row 0 = initial
 * state), then once at the end "mu" (row r: q x d), "lambda" (row r: nl x d x d), "plogLik" and "Ychange" (row i: 1).
 * `dims` is the per-row shape mcmcChain.R stores as the dataset's "dims" attribute.  With a sink registered z /
 * weights / Y may be NULL in the *_out structs (BASELINE config C3: 7.2 GB of Y snapshots never sit in host memory).
 * A non-zero return of write() fails the call with BSB200_ERR_BUFFER.  One sink per process; NULL unregisters. */
typedef struct bsb200_sink {
  int (*write)(void *user, const char *param, int64_t row, const void *data, int64_t count, int32_t is_f64,
               const int64_t *dims, int32_t ndims);
  void *user;
} bsb200_sink;
void bsb200_set_snapshot_sink(const bsb200_sink *sink);
/* Built-in sink: a directory with one raw little-endian file per parameter ("<param>.bin", one row per kept
 * iteration) and "meta.json" (dtype, rows, cols, dims) -- the layout of mcmcChain.R's HDF5 datasets without libhdf5
 * (reader: INTEGRATION.md).  open registers the sink, close writes meta.json and unregisters it. */
int bsb200_chain_file_open(const char *dir);
int bsb200_chain_file_close(void);

/* ------------------------------------------------------------------------------------------------
 * Neighbour construction (host).  Two-pass sizing: call with idx == NULL (or too small a capacity)
 * to obtain *nnz, then again with a buffer of that many entries.
 * ---------------------------------------------------------------------------------------------- */

/* .find_neighbors (R/spatialCluster.R:227-302): lattice neighbours from integer array coordinates;
 * hex offsets (+-2,0),(+-1,+-1) or square offsets (0,+-1),(+-1,0); ascending by spot index. */
int bsb200_neighbors_lattice(const int32_t *array_row, const int32_t *array_col, int64_t n,
                             int32_t platform, int64_t *ptr, int32_t *idx, int64_t cap, int64_t *nnz);

/* find_neighbors (R/utils.R:13-26): 0 < dist <= radius, ascending.  positions: n x 2 col-major. */
int bsb200_neighbors_radius(const double *positions, int64_t n, double radius, int32_t metric,
                            int64_t *ptr, int32_t *idx, int64_t cap, int64_t *nnz);

/* convert_neighbors (src/cluster.cpp:134-145) + str_split (src/utils.h:10-62) + Neighbor ctor
 * (src/neighbor.h:67-80): n0 strings "a,b,c" with 1-based ids, NULL = NA -> 0-based CSR. */
int bsb200_parse_neighbor_strings(const char *const *strings, int64_t n0, int64_t *ptr, int32_t *idx,
                                  int64_t cap, int64_t *nnz);

/* find_subspot_neighbors (src/cluster.cpp:161-215): candidates = all subspots of (neighbour spots
 * ascending, then the own spot); kept when 1e-6 < Manhattan distance < dist; candidate order is
 * preserved.  positions: n x 2 col-major, subspot s = r*n0 + j0. */
int bsb200_subspot_neighbors(const double *positions, int64_t n, int64_t n0, const int64_t *spot_ptr,
                             const int32_t *spot_idx, double dist, int64_t *ptr, int32_t *idx,
                             int64_t cap, int64_t *nnz);

/* convert_neighbors2mtx (src/cluster.cpp:147-159): n x 4 col-major, 1-based, zero padded. */
int bsb200_neighbors_to_matrix4(const int64_t *ptr, const int32_t *idx, int64_t n, int64_t *out);

/* Graph colouring used by the chromatic scan (no reference counterpart: the reference scans
 * sequentially, src/cluster.cpp:275).  Greedy in index order on the symmetrised graph (self loops
 * ignored), so that no spot ever reads the label of a spot of its own colour. */
int bsb200_colour_graph(const int64_t *ptr, const int32_t *idx, int64_t n, int32_t *colour,
                        int32_t *ncolours);

/* Closed-form lattice colourings: square (row + col) mod 2; hex
 * ((col - 3 row) / 2) mod 3, which needs col == row (mod 2) as on Visium. */
int bsb200_colour_lattice(const int32_t *array_row, const int32_t *array_col, int64_t n,
                          int32_t platform, int32_t *colour, int32_t *ncolours);

/* ------------------------------------------------------------------------------------------------
 * Cluster samplers: iterate / iterate_vvv / iterate_t / iterate_t_vvv
 * (src/cluster.cpp:217-321 / 323-444 / 446-568 / 570-710; R/RcppExports.R:4-18).
 * (model, precision) selects the reference function exactly as cluster() does
 * (R/spatialCluster.R:94-106).
 * ---------------------------------------------------------------------------------------------- */
typedef struct bsb200_cluster_args {
  int32_t model;     /* BSB200_MODEL_*      */
  int32_t precision; /* BSB200_PRECISION_*  */
  int64_t n;         /* spots               */
  int32_t d;         /* PCs, 1..32          */
  int32_t q;         /* clusters, 2..255    */
  int32_t nrep;      /* loop runs i = 1 .. nrep-1 */
  int32_t thin;
  double gamma, alpha, beta;
  const double *Y;         /* n x d col-major                                   */
  const int64_t *nbr_ptr;  /* df_j as CSR (0-based ids)                          */
  const int32_t *nbr_idx;
  const int32_t *init;     /* n labels, 1-based                                  */
  const double *mu0;       /* d                                                  */
  const double *lambda0;   /* d x d                                              */
  uint64_t seed;           /* Philox key; the R shim draws it from R's RNG      */
  uint32_t compat;         /* BSB200_COMPAT_* bitmask                            */
  int32_t device;          /* CUDA device ordinal                                */
  const int32_t *colour;   /* optional n colours (0-based) from bsb200_colour_graph, or NULL */
  int32_t ncolours;
  int32_t groups;          /* virtual shards of the statistics fold: 0 = automatic (1, or 64 for n >= 262,144
                              or world > 1); chains are bit-identical across GPU counts for equal `groups` */
  /* spot sharding of ONE chain over `world` processes, one GPU each (every process passes the full inputs) */
  int32_t world;           /* 0 or 1: single GPU */
  int32_t rank;
  const void *comm_id;     /* 128 bytes from bsb200_comm_unique_id() of rank 0, identical on all ranks */
} bsb200_cluster_args;

/* nsnap = nrep/thin + 1 rows.  Any pointer may be NULL to skip that output. */
typedef struct bsb200_cluster_out {
  int32_t *z;       /* nsnap x n, 1-based                                               */
  double *mu;       /* nsnap x (q*d), row-wise vectorised (mu_1,1..mu_1,d, mu_2,1..)    */
  double *lambda;   /* nsnap x nl x (d*d) col-major; nl = 1 (equal) or q (variable)     */
  double *weights;  /* nsnap x n (t model only)                                         */
  double *plogLik;  /* nrep, [0] = NaN                                                  */
  int32_t rows_done; /* out: snapshot rows written (== nsnap unless cancelled)          */
  int32_t ncolours;  /* out: colour classes used by the chromatic scan                  */
  double setup_ms;   /* out: host->device staging + graph setup, wall clock             */
  double device_ms;  /* out: device time of the sampling loop (CUDA events)             */
  int64_t own_lo, own_hi; /* out: spots [own_lo, own_hi) of z / weights were written by this rank
                             (all of them when world <= 1); mu / lambda / plogLik are complete on every rank.
                             A rank of a sharded chain writes row 0 and its own columns of the later rows and leaves
                             the other ranks' columns as the caller passed them (no pass over the full matrices) */
  /* Optional device-side post-sampling reduction: the per-spot MODE of the label snapshots of rows
   * r >= mode_from, with R's Mode() tie rule (R/utils.R:63-66: among equally frequent labels the one that appears
   * first in the kept rows) -- what spatialCluster() computes with apply(zs, 2, Mode) (R/spatialCluster.R:204-213,
   * mode_from = burn.in %/% thin + 1).  n int32, 1-based; 0 where no row was kept.  With z == NULL the label
   * snapshots never leave the GPU (440 MB per 11 rows at 10M spots). */
  int32_t *z_mode;
  int32_t mode_from;
} bsb200_cluster_out;

int bsb200_cluster(const bsb200_cluster_args *args, bsb200_cluster_out *out);

/* Multi-GPU (one process per GPU of one box): rank 0 creates the communicator id, the host layer hands the 128
 * bytes to every rank (torch.distributed broadcast in bench.py), each rank passes them in
 * bsb200_cluster_args.comm_id; ONE id per sharded chain / bsb200_cluster call.  Per iteration the ranks exchange
 * the labels of boundary spots after every colour phase and the per-shard statistics once.  The id selects the
 * transport: by default the library's own kernels store into peer GPU memory over NVLink (CUDA IPC mappings,
 * epoch flags, host rendezvous in POSIX shared memory); with BSB200_COMM=nccl in rank 0's environment when the
 * id is made, NCCL send/recv + allgather (libnccl.so.2 bound at run time).  Both give the same bits.
 * Environment knobs: BSB200_PEER_TIMEOUT (s, default 20: how long a kernel waits for a peer's flag before the
 * chain fails with BSB200_ERR_CUDA), BSB200_RENDEZVOUS_TIMEOUT (s, default 120: host-side wait for the other
 * ranks at set-up).  There is no reference counterpart (the reference is single-process). */
int bsb200_comm_unique_id(void *id128);
/* Host-only view of a rank's share of a sharded chain (own range, ghost spots, halo lists); for tests. */
int bsb200_shard_plan(const int64_t *ptr, const int32_t *idx, int64_t n, const int32_t *colour, int32_t ncolours,
                      int32_t groups, int32_t world, int32_t rank, int64_t *lo, int64_t *hi, int32_t *ghosts,
                      int64_t cap, int64_t *nghosts, int32_t *counts, int32_t *send_ids, int32_t *recv_ids);

/* Resident-chain interface: the same sampler with state kept in HBM between calls.  bench.py uses
 * it to time sweeps with inputs already on the device; bsb200_cluster() is create + run + destroy. */
typedef struct bsb200_chain bsb200_chain;
int bsb200_chain_create(const bsb200_cluster_args *args, bsb200_chain **chain);
/* Advances `iters` iterations (continuing the iteration counter); device time in *device_ms. */
int bsb200_chain_run(bsb200_chain *chain, int32_t iters, double *device_ms);
/* As bsb200_chain_run but launches kernel by kernel with CUDA events around every z-sweep launch:
 * sweep_ms = total time inside the sweep kernels, launches = number of sweep launches timed, param_ms = statistics
 * fold (+ allgather) + k_param, halo_ms = boundary-label exchanges (0 on a single GPU). */
int bsb200_chain_profile(bsb200_chain *chain, int32_t iters, double *sweep_ms, int64_t *launches,
                         double *param_ms, double *halo_ms);
/* Current state in the caller's (original) spot order; NULL pointers are skipped.
 * plogLik receives the values of all iterations run so far (length = iterations done + 1). */
int bsb200_chain_state(bsb200_chain *chain, int32_t *z, double *weights, double *mu, double *lambda,
                       double *plogLik, int64_t *iters_done);
int bsb200_chain_destroy(bsb200_chain *chain);
/* Number of kernels this library has launched in this process (for bench.py's gpu_launches). */
int64_t bsb200_kernel_launches(void);
/* Device memory of destroyed chains stays in the device's stream-ordered pool for the next chain (allocation and
 * release of the ~3 GB a 10M-spot chain holds otherwise cost up to hundreds of milliseconds per call), as do the
 * page-locked snapshot buffers.  This hands both back to the driver; nothing else in the library depends on it. */
int bsb200_release_memory(int32_t device);

/* ------------------------------------------------------------------------------------------------
 * State-level parity probes (evaluate one quantity of the sampler on a caller-supplied state).
 * ---------------------------------------------------------------------------------------------- */

/* Spot x cluster log-density, n x q col-major:
 * form 0: dmvnrm_arma_fast(y_j, mu_k, sigma_k / w_j) (src/cluster.cpp:83-106, Q1 under compat);
 * form 1: dmvnrm_prec_arma_fast(y_j, mu_k, mat_k * w_j) (:109-132).
 * mu: q x d row-major; mat: nl x d x d col-major (nl = 1 or q); w: n or NULL. */
int bsb200_eval_loglik(int32_t device, int32_t form, uint32_t compat, const double *Y, int64_t n,
                       int32_t d, int32_t q, const double *mu, const double *mat, int32_t nl,
                       const double *w, double *out);

/* Sufficient statistics the mu / lambda updates consume (src/cluster.cpp:248-265, 482-500):
 * nk[q] = sum_j w_j [z_j = k], sk[q x d] row-major = sum w_j y_j, scatter[nl x d x d] =
 * sum w_j (y_j - mu_{z_j})(y_j - mu_{z_j})^T (pooled when nl == 1). */
int bsb200_eval_suffstats(int32_t device, const double *Y, int64_t n, int32_t d, int32_t q,
                          const int32_t *z, const double *w, const double *mu, int32_t nl,
                          double *nk, double *sk, double *scatter);

/* One dry sweep on the chain's current state: for every spot the proposal the chain would make at
 * its next iteration and the two values the MH step compares (src/cluster.cpp:283-298), without
 * changing the state.  All outputs are length n, original spot order. */
int bsb200_chain_dry_sweep(bsb200_chain *chain, int32_t *z_new, double *w_new, double *h_prev,
                           double *h_new, double *prob, int32_t *accepted);

/* Test hook for the host rendezvous of the peer transport (no GPU involved): `rounds` allgathers of `bytes` bytes
 * per rank through the shared-memory segment named by a peer-transport id; in round r rank p contributes mine[i] + r.
 * all_last (world * bytes) receives the last round. */
int bsb200_debug_rendezvous(const void *id128, int32_t world, int32_t rank, const uint8_t *mine, int64_t bytes,
                            int32_t rounds, uint8_t *all_last);

/* Keyed random variates exactly as the kernels draw them (u in (0,1), N(0,1), Gamma(shape, scale)):
 * kind 0 uniform pair (count must be 2), 1 normals m0.., 2 gammas for index0.. */
int bsb200_debug_rng(int32_t device, int32_t kind, uint64_t seed, uint32_t purpose, uint32_t index,
                     uint32_t iter, uint32_t slot, double shape, double scale, int32_t count,
                     double *out);

/* ------------------------------------------------------------------------------------------------
 * Initial labels: .init_cluster (R/spatialCluster.R:317-328), the step before the samplers
 * ---------------------------------------------------------------------------------------------- */
enum { BSB200_INIT_KMEANS = 0, BSB200_INIT_MCLUST = 1 };
/* init.method "kmeans": k-means++ seeding + Lloyd iterations; "mclust": EM of the Gaussian mixture with one shared
 * covariance (mclust's "EEE" model) started from that partition; labels = argmax of the responsibilities.
 * stats::kmeans / mclust::Mclust draw from R's RNG and use algorithms of their own, so their labels are not
 * reproducible bit for bit; the contract kept is: n labels in 1..q, every label present, deterministic under `seed`
 * (Philox key, as for the samplers).  Y: n x d col-major.  max_iter <= 0: 100.  centers (optional): q x d row-major.
 * Runs on the device: at 10M spots this R step would otherwise dominate spatialCluster(). */
int bsb200_init_cluster(int32_t device, const double *Y, int64_t n, int32_t d, int32_t q, int32_t method,
                        uint64_t seed, int32_t max_iter, int32_t *init, double *centers, int32_t *iters_done);

/* ------------------------------------------------------------------------------------------------
 * Helpers of spatialEnhance's benchmarking code (src/compare_enhanced.cpp; R/spatialEnhance.R:570-585)
 * ---------------------------------------------------------------------------------------------- */

/* map_subspot2ref (src/compare_enhanced.cpp:47-92): for every subspot the reference spot(s) at minimal squared
 * Euclidean distance; exact ties are all reported, as "find(dists == min(dists))" does.  Coordinates are ns x k and
 * nr x k col-major (k <= 8).  The reference returns a 3-column matrix with one row (i, j, min dist^2) per tie; here:
 * row_ptr[ns + 1] offsets of subspot i's ties in ref_idx (ascending j), min_dist2[ns], *nrows = total rows.
 * Two-pass sizing like the neighbour builders: ref_idx == NULL returns *nrows only. */
int bsb200_map_subspot2ref(int32_t device, const double *subspot_coords, int64_t ns, const double *ref_coords,
                           int64_t nr, int32_t k, int64_t *row_ptr, int32_t *ref_idx, int64_t cap, double *min_dist2,
                           int64_t *nrows);

/* compute_corr (src/compare_enhanced.cpp:94-132): cor[i] = Pearson correlation of row i of m1 and m2
 * (both nrow x ncol col-major).  The reference's first output column is the row index i itself. */
int bsb200_compute_corr(int32_t device, const double *m1, const double *m2, int64_t nrow, int64_t ncol, double *cor);

/* ------------------------------------------------------------------------------------------------
 * iterate_deconv (src/cluster.cpp:742-1141; R/RcppExports.R:20-22)
 * ---------------------------------------------------------------------------------------------- */
typedef struct bsb200_deconv_args {
  const double *subspot_positions; /* n x 2 col-major                                   */
  double dist;
  const int64_t *spot_nbr_ptr;     /* parsed spot.neighbors (bsb200_parse_neighbor_strings) */
  const int32_t *spot_nbr_idx;
  const double *Y;                 /* n x d col-major, parent PCs replicated; NOT mutated (Q10) */
  int32_t tdist;
  int32_t nrep, thin;
  int64_t n, n0;
  int32_t d;
  double gamma;
  int32_t q;
  const int32_t *init;             /* n labels, 1-based */
  int32_t subspots;
  int32_t verbose;
  double jitter_scale;             /* 0 => adaptive (Robust Adaptive Metropolis) */
  int32_t adapt_before;
  double c;
  const double *mu0;
  const double *lambda0;
  double alpha, beta;
  int32_t thread_num;              /* accepted for signature parity; ignored */
  uint64_t seed;
  int32_t device;
  int32_t reserved;
} bsb200_deconv_args;

typedef struct bsb200_deconv_out {
  int32_t *z;       /* nsnap x n                                          */
  double *mu;       /* nsnap x (q*d)                                      */
  double *lambda;   /* nsnap x (d*d)                                      */
  double *weights;  /* nsnap x n (ones for the normal model)              */
  double *Y;        /* nsnap x (n x d col-major), or NULL                 */
  double *Y_mean;   /* optional running mean of the Y snapshots r >= mean_from (n x d col-major) */
  int32_t mean_from;
  int32_t mode_from; /* rows r >= mode_from enter z_mode (below)                                  */
  int32_t *z_mode;   /* optional per-subspot mode of the label snapshots, R's Mode() tie rule
                        (R/spatialEnhance.R:450-458, R/utils.R:63-66); n int32, 1-based            */
  double *Ychange;  /* nrep, [0] = NaN                                    */
  double *plogLik;  /* nrep, [0] = NaN                                    */
  int64_t *df_j;    /* n x 4 col-major, 1-based zero padded               */
  int32_t rows_done;
  int32_t ncolours;
  double setup_ms, device_ms;
} bsb200_deconv_out;

int bsb200_deconv(const bsb200_deconv_args *args, bsb200_deconv_out *out);

#ifdef __cplusplus
}
#endif
#endif /* BSB200_H */
