/*
 * vbsky_b200.h -- C ABI of the B200-native VBSKY hot path (libvbsky_b200.so).
 *
 * Drop-in boundary for the ELBO's data-parallel path of jthlab/vbsky: Felsenstein pruning
 * log-likelihood + gradient over (trees x MC samples x site patterns), the BDSKY tree prior
 * density + gradient, and the height-transform chain between them.  The reference has no FFI
 * of its own (it is pure Python on JAX); every entry point below cites the reference call it
 * replaces (file:line under /root/reference/vbsky).  INTEGRATION.md shows the jax.ffi /
 * ctypes bindings.
 *
 * Conventions
 *   - plain pointers and sizes; no torch / jax types.  "dev" pointers are CUDA device pointers
 *     in the current device's primary context, "host" pointers are ordinary host memory.
 *   - every call returns 0 on success or a negative VBSKY_E* code; vbsky_last_error() returns
 *     the calling thread's last message.  Nothing throws across the ABI.
 *   - device work is enqueued on the given stream (a cudaStream_t passed as void*) and is
 *     asynchronous unless the name ends in _host; no global mutable state; handles are
 *     immutable after creation and may be shared by threads.
 *   - arrays are row-major; node indices int32; reals are fp64 (dtype "f64").
 *   - node numbering is the reference's (tree_data.py:464-507): tips 0..n-1, internal nodes
 *     n..2n-2 with children < parent, root = 2n-2.
 *   - a "unit" is one (tree, Monte-Carlo sample) pair; unit u uses tree unit_tree[u].
 */
#ifndef VBSKY_B200_H
#define VBSKY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VBSKY_OK 0
#define VBSKY_EINVAL (-1)   /* bad argument / shape / malformed tree */
#define VBSKY_ECUDA (-2)    /* CUDA runtime error (message has the cudaError string) */
#define VBSKY_ENOMEM (-3)   /* workspace too small / allocation failure */
#define VBSKY_EKEY (-4)     /* unknown sequence character (reference: KeyError, substitution.py:48) */

const char* vbsky_last_error(void);
int vbsky_version(void);
/* Number of CUDA devices visible; <0 on error.  The product path has no CPU fallback. */
int vbsky_device_count(void);

/* ------------------------------------------------------------------------------------------
 * Substitution model: mirrors SubstitutionModel(pi, P, d, Pinv) (substitution.py:51-55).
 * expm(t) = P diag(exp(t d)) Pinv (substitution.py:61-62), Q = P diag(d) Pinv (:57-59).      */
typedef struct vbsky_subst_model {
  double pi[4];
  double P[16];
  double d[4];
  double Pinv[16];
} vbsky_subst_model;

/* HKY(kappa) with uniform base frequencies as substitution.py:73-87 builds it from
 * msprime.HKY: off-diagonal rates 1/(2+kappa) (transversion), kappa/(2+kappa) (transition),
 * diagonal -1.  Closed-form spectrum (no LAPACK); the degenerate eigen-pair's basis differs
 * from eigh's but P f(d) Pinv is identical.  JC69 == HKY(1) (substitution.py:90-92).          */
int vbsky_hky(double kappa, vbsky_subst_model* out);
/* Q = P diag(d) Pinv into q[16]. */
int vbsky_subst_rate_matrix(const vbsky_subst_model* m, double* q);
/* d(d)/d(kappa) for vbsky_hky's spectrum into dd[4].  The eigenvectors of HKY with uniform
 * frequencies do not depend on kappa, so dQ/dkappa = P diag(dd) Pinv commutes with Q.        */
int vbsky_hky_dkappa(double kappa, double* dd);

/* ------------------------------------------------------------------------------------------
 * Tip encoding and site-pattern compression (host, bit-exact).
 * Codes are 4-bit state masks: bit0=A bit1=C bit2=G bit3=T, i.e. the {0,1}^4 rows of
 * PARTIAL_DICT (substitution.py:10-35) packed into one byte.                                 */

/* encode_partials (substitution.py:38-48) for n sequences of length L (seqs[i*L + l]):
 * writes codes[l*n + i] (site-major, tips in the given order).  Unknown or lower-case
 * characters give VBSKY_EKEY like the reference's KeyError.                                  */
int vbsky_encode_tips(const char* seqs, int64_t n, int64_t L, uint8_t* codes);

/* np.unique(tip_partials, axis=0, return_counts=True) of fasta.py:547 on packed codes:
 * rows of codes[L][n] sorted lexicographically (by the unpacked [n,4] 0/1 rows, which is the
 * order NumPy gives), duplicates merged.  Optional perm[n] first permutes the columns (tips)
 * like tip_partials0[:, perm] (fasta.py:545-546).  Outputs patterns[S][n], counts[S]; both
 * buffers must hold L rows.                                                                  */
int vbsky_compress_patterns(const uint8_t* codes, int64_t L, int64_t n, const int64_t* perm,
                            uint8_t* patterns, int64_t* counts, int64_t* S_out);

/* Device versions (one CTA per alignment; same results bit for bit): seqs dev [n][L] chars ->
 * codes dev [L][n]; codes dev [T][L][n] (+ optional perm dev [T][n]) -> patterns dev [T][L][n]
 * (first S_out[t] rows valid), counts dev [T][L], S_out dev [T].  The radix sort is a stable
 * counting sort per tip column on the bit-reversed masks, least significant column first.
 * vbsky_encode_tips_device synchronises the stream (it reports the KeyError position).          */
int vbsky_encode_tips_device(const char* seqs, int64_t n, int64_t L, uint8_t* codes,
                             int64_t* bad_index, void* stream);
size_t vbsky_compress_workspace_bytes(int32_t T, int64_t L);
int vbsky_compress_patterns_device(int32_t T, int64_t L, int32_t n, const uint8_t* codes,
                                   const int32_t* perm, uint8_t* patterns, int64_t* counts,
                                   int32_t* S_out, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Starting trees (SeqData.sample_trees, fasta.py:203-296), batched over the T subsamples.
 *
 * vbsky_snp_pack_device: seqs dev [N][L] chars -> planes dev [N][4][ceil(L/32)] bit planes
 * (A, C, G, T after upper-casing; everything else is "ignored"), once per alignment.
 * vbsky_snp_dists_device: what `snp-dists -b` prints for the subsample rows[t][0..n) of the
 * alignment (fasta.py:268-281): dist dev [T][n][n] int32, dist[r][c] = #columns where the two
 * characters differ and both are in AGTC.
 * vbsky_supgma_distances_device: get_distance_matrix (upgma.py:25-62) on dist's lower triangle
 * and times dev [T][n]: omega[t] = last coefficient of the least-squares fit (one intercept
 * per time class of the later sample, or a single intercept when single_theta), and
 * D dev [T][n][n] (symmetric, zero diagonal) = y + omega (t_r + t_c - 2 min t).
 * vbsky_upgma_device: Bio.Phylo's DistanceTreeConstructor.upgma (fasta.py:283) on D (which is
 * overwritten): merges dev [T][n-1][2] = (clade1, clade2) of the k-th inner clade in child
 * order, a clade being a tip position < n or n + (index of an earlier merge); heights dev
 * [T][n-1] = min_dist / 2.  Ties are broken as Biopython's scan does (last `>=` hit); the new
 * row is the plain average of the two merged rows.  n <= 4096.                                */
size_t vbsky_snp_planes_bytes(int64_t N, int64_t L);
int vbsky_snp_pack_device(const char* seqs, int64_t N, int64_t L, uint32_t* planes, void* stream);
int vbsky_snp_dists_device(const uint32_t* planes, int64_t N, int64_t L, const int32_t* rows,
                           int32_t T, int32_t n, int32_t* dist, void* stream);
int vbsky_supgma_distances_device(int32_t T, int32_t n, const int32_t* dist, const double* times,
                                  int32_t single_theta, double* D, double* omega, void* stream);
int vbsky_upgma_device(int32_t T, int32_t n, double* D, int32_t* merges, double* heights,
                       void* stream);

/* ------------------------------------------------------------------------------------------
 * Tree layout: an ensemble of T trees with n tips each, given as the reference's TreeData
 * arrays (tree_data.py:17-23).  Builds, per tree, the evaluation schedules of the pruning
 * kernel (a depth-first order that keeps all live partial vectors on chip), the level order
 * of the height transform, siblings (tree_data.py:68-78) and lower sampling times
 * (tree_data.py:80-117), and uploads them to the current device.
 * Validates: children < parent for internal nodes, root == 2n-2, child_parent[root] == -1,
 * parent_child consistent with child_parent.                                                 */
typedef struct vbsky_layout vbsky_layout;

int vbsky_layout_create(int32_t T, int32_t n,
                        const int32_t* parent_child /* host [T][2n-1][2] */,
                        const int32_t* child_parent /* host [T][2n-1]    */,
                        const int32_t* postorder    /* host [T][2n-1]    */,
                        const double* sample_times  /* host [T][n]       */,
                        vbsky_layout** out);
void vbsky_layout_destroy(vbsky_layout* lay);
/* info[0]=T, [1]=n, [2]=max live vectors in the up pass, [3]=same for the down pass,
 * [4]=max number of levels of the height transform. */
int vbsky_layout_info(const vbsky_layout* lay, int32_t info[8]);
/* Host copies for inspection / tests: up[n-1][4], down[n-1][4] (encoding in DESIGN.md),
 * down_nodes[2n-2] (node whose branch gradient each down-pass output slot holds),
 * lower_sampling_times[2n-1], siblings[2n-1].  Any output may be NULL.                       */
int vbsky_layout_get(const vbsky_layout* lay, int32_t t, int32_t* up, int32_t* down,
                     int32_t* down_nodes, double* lower_sampling_times, int32_t* siblings);
/* The down pass's ring-stage chunking (inspection / tests): n_chunks, chunk_start[cap+1] (first step of each chunk;
 * chunks hold <= 3 consecutive steps and never split a step from the cherry children it rebuilds), msg_start[cap+1]
 * (prefix of msg_rows per chunk), n_msgs and msg_rows[max(n-2,1)] (scratch rows of the internal non-cherry children in
 * consumption order).  cap = vbsky_layout_info()[5].  Any output may be NULL.                                    */
int vbsky_layout_get_chunks(const vbsky_layout* lay, int32_t t, int32_t* n_chunks, int32_t* chunk_start,
                            int32_t* msg_start, int32_t* n_msgs, int32_t* msg_rows);

/* ------------------------------------------------------------------------------------------
 * Tip data of the ensemble on the device: TipData(partials[S,n,4], counts[S]) per tree
 * (util.py:13-15) after pad_tips (fasta.py:325-331), as packed codes.
 * patterns: host [T][S][n] 4-bit masks (row s of tree t = pattern s over tips in node order);
 * counts:   host [T][S] pattern multiplicities (0 for padding rows).
 * The codes are stored in the order the layout's schedules consume them, so tips belong to the
 * layout they were created with (T and n are taken from it).                                  */
typedef struct vbsky_tips vbsky_tips;

int vbsky_tips_create(const vbsky_layout* lay, int64_t S, const uint8_t* patterns,
                      const double* counts, vbsky_tips** out);
void vbsky_tips_destroy(vbsky_tips* tips);
/* Replace the pattern weights (host [T][S]) of an existing tips object; synchronous.  Used to obtain
 * d/d bl of sum_s w[s] * site_ll[s] for arbitrary cotangents w (the vmapped prune_loglik's transpose). */
int vbsky_tips_set_counts(vbsky_tips* tips, const double* counts);

/* ------------------------------------------------------------------------------------------
 * Pruning likelihood and gradient.  Replaces, for all units at once,
 *     vmap(prune.prune_loglik, (None,None,0,...))(bl, Q, partials, td) * counts   (bdsky.py:331-343)
 * and its custom derivative prune_loglik_jvp (prune.py:134-150, eq. 9).
 *
 *   unit_tree  dev  [U]            tree index of each unit
 *   bl         dev  [U][2n-2]      branch length above each node, ALREADY times clock_rate
 *   site_ll    dev  [U][S]         per-pattern log-likelihood (prune_loglik's value); may be NULL
 *   ll         dev  [U]            sum_s counts[s] * site_ll[s]
 *   dll_dbl    dev  [U][2n-2]      d ll / d bl; NULL => value only (no scratch traffic at all)
 *   ws         dev                 workspace of >= vbsky_prune_workspace_bytes(...) bytes
 * Gradients w.r.t. Q, pi and the tip partials are not produced (prune.py:149).                */
size_t vbsky_prune_workspace_bytes(const vbsky_layout* lay, const vbsky_tips* tips, int32_t U,
                                   int32_t want_grad);
int vbsky_prune_value_and_grad(const vbsky_layout* lay, const vbsky_tips* tips, int32_t U,
                               const int32_t* unit_tree, const double* bl,
                               const vbsky_subst_model* model, double* site_ll, double* ll,
                               double* dll_dbl, void* ws, size_t ws_bytes, void* stream);

/* Same call with fp32 messages / tables / scratch (half the HBM traffic, FP32 pipe); inputs and
 * outputs stay fp64, P(t) is built in fp64 and rounded once, the final log and all sums over
 * patterns are fp64.  Accuracy target of BASELINE.json: 1e-5 relative on log-likelihoods.      */
size_t vbsky_prune_workspace_bytes_f32(const vbsky_layout* lay, const vbsky_tips* tips, int32_t U,
                                       int32_t want_grad);
int vbsky_prune_value_and_grad_f32(const vbsky_layout* lay, const vbsky_tips* tips, int32_t U,
                                   const int32_t* unit_tree, const double* bl,
                                   const vbsky_subst_model* model, double* site_ll, double* ll,
                                   double* dll_dbl, void* ws, size_t ws_bytes, void* stream);

/* Gradient with respect to substitution parameters (asked for by the north star; the reference
 * drops these tangents, prune.py:149).  Same evaluation as vbsky_prune_value_and_grad, but the
 * quadratic form of eq. (9) uses G = P diag(g) Pinv instead of Q = P diag(d) Pinv:
 *   out[u][b] = sum_s counts[s] * (pre_b^T G post_b) / L_s          (g = d gives dll_dbl).
 * For a parameter theta that moves only the spectrum (dQ/dtheta = P diag(dd/dtheta) Pinv, e.g.
 * kappa of vbsky_hky via vbsky_hky_dkappa), dP_b/dtheta = bl_b G P_b and therefore
 *   d ll[u] / d theta = sum_b bl[u][b] * out[u][b].                                           */
int vbsky_prune_spectral_grad(const vbsky_layout* lay, const vbsky_tips* tips, int32_t U,
                              const int32_t* unit_tree, const double* bl,
                              const vbsky_subst_model* model, const double* g, double* ll,
                              double* out, void* ws, size_t ws_bytes, void* stream);

/* Gradient w.r.t. an ARBITRARY parameter theta of the substitution model (north star item 3; the reference drops
 * these tangents, prune.py:149).  dQ [16] = dQ/dtheta (row-major; any direction, need not commute with Q: base
 * frequencies of HKY/GTR, GTR exchangeabilities).  One evaluation of the pruning kernel with a per-branch matrix in
 * eq. (9)'s quadratic form, D_b = dP(t_b)/dtheta P(t_b)^-1:
 *   out[u][b] = sum_s counts[s] * (w^T D_b m_b) / (w^T m_b)   -- branch b's contribution,
 *   d ll[u] / d theta (Q-part) = sum_b out[u][b]              (no branch-length factor),
 * dpi [U][4] (may be NULL) = d ll[u] / d pi_i at fixed Q (the root-frequency term of prune.py:131); the caller adds
 * sum_i dpi[u][i] * d pi_i / d theta.  Workspace as vbsky_prune_workspace_bytes(..., want_grad = 1).  fp64 only.   */
int vbsky_prune_param_grad(const vbsky_layout* lay, const vbsky_tips* tips, int32_t U,
                           const int32_t* unit_tree, const double* bl,
                           const vbsky_subst_model* model, const double* dQ, double* ll,
                           double* out, double* dpi, void* ws, size_t ws_bytes, void* stream);

/* Counters of the last prune call made by this thread: out[0]=kernel launches,
 * out[1]=persistent CTAs, out[2]=sites per CTA tile, out[3]=dynamic smem bytes per CTA,
 * out[4]=algorithmic HBM bytes of the fused pruning kernel (DESIGN.md formula).               */
int vbsky_prune_last_stats(int64_t out[8]);

/* ------------------------------------------------------------------------------------------
 * Model parameters of bdsky.loglik (bdsky.py:268-345), one row per unit, all device pointers:
 * the reference's `params` dict after optim.unpack (optim.py:19-32), vmapped over the sample
 * axis (optim.py:72).  proportions [U][n-2]; root_proportion, origin, origin_start, clock_rate
 * [U]; R, delta, s, rho [U][m]; grid [U][m+1] (optional, see below).  Zero-initialise the struct: fields may be
 * appended.                                                                                   */
typedef struct vbsky_params {
  const double* proportions;
  const double* root_proportion;
  const double* origin;
  const double* origin_start;
  const double* clock_rate;
  const double* R;
  const double* delta;
  const double* s;
  const double* rho;
  /* Prespecified skyline grid (bdsky.py:317-319, equidistant_intervals=False): [U][m+1] or NULL.  When given,
   * times = tm - grid with times[0] = 0 instead of linspace(0, tm, m+1).                                       */
  const double* grid;
} vbsky_params;

/* Gradients w.r.t. the same fields (origin_start shares origin's gradient; both enter as a sum).
 * R/delta/s/rho may be NULL when only the height chain is differentiated.                    */
typedef struct vbsky_param_grads {
  double* proportions;
  double* root_proportion;
  double* origin;
  double* clock_rate;
  double* R;
  double* delta;
  double* s;
  double* rho;
  double* grid; /* [U][m+1] or NULL; entry 0 is always 0 (times[0] is pinned to 0) */
} vbsky_param_grads;

/* TreeData.height_transform (tree_data.py:287-367) and the glue of bdsky.loglik
 * (bdsky.py:292-322) for all units: heights [U][2n-1]; bl_scaled [U][2n-2] = (h[parent]-h) *
 * clock_rate; optional xs [U][2n-1] = tm - heights, times [U][m+1] = linspace(0, tm, m+1),
 * lam/psi/mu [U][m] (_bdsky_transform, bdsky.py:219-227) with tm = origin + origin_start.
 * Pass times == NULL (and m == 0) to skip the skyline outputs.                                */
int vbsky_heights_forward(const vbsky_layout* lay, int32_t U, const int32_t* unit_tree, int32_t m,
                          const vbsky_params* in, double* heights, double* bl_scaled, double* xs,
                          double* times, double* lam, double* psi, double* mu, void* stream);

/* Adjoint of vbsky_heights_forward: given d/d(bl_scaled), d/d(xs), d/d(times), d/d(lam, psi, mu)
 * (any may be NULL = 0) accumulate the gradient w.r.t. proportions, root_proportion, origin,
 * clock_rate and (if out->R != NULL) R, delta, s.  ws: >= U*(2n-1)*8 bytes.                   */
int vbsky_heights_backward(const vbsky_layout* lay, int32_t U, const int32_t* unit_tree, int32_t m,
                           const vbsky_params* in, const double* heights,
                           const double* d_bl_scaled, const double* d_xs, const double* d_times,
                           const double* d_lam, const double* d_psi, const double* d_mu,
                           vbsky_param_grads* out, void* ws, size_t ws_bytes, void* stream);

/* _tree_loglik (bdsky.py:18-216): birth-death-skyline log-density of Stadler et al. (2013),
 * Thm 1, for U units at once, and its gradient (the reference obtains it by JAX autodiff).
 * lam, psi, mu, rho [U][m]; xs [U][2n-1] (forward node times, tips first); times [U][m+1].
 * Out-of-range interval indices clamp (jnp.take clip semantics).  Outputs: lp [U];
 * if dlam != NULL also dlam, dpsi, dmu, drho [U][m], dxs [U][2n-1], dtimes [U][m+1].           */
size_t vbsky_bdsky_workspace_bytes(int32_t U, int32_t n);
int vbsky_bdsky_value_and_grad(int32_t U, int32_t n, int32_t m, const double* lam,
                               const double* psi, const double* mu, const double* rho,
                               const double* xs, const double* times,
                               int32_t condition_on_survival, double* lp, double* dlam,
                               double* dpsi, double* dmu, double* drho, double* dxs,
                               double* dtimes, void* ws, size_t ws_bytes, void* stream);

/* bdsky.loglik (bdsky.py:268-345) without the user prior term c[0]: for every unit
 *   ll = [flags&1] _tree_loglik(...) + [flags&2] sum_s counts[s] * prune_loglik(...)
 * and the gradient w.r.t. every field of vbsky_params (grads == NULL => value only).
 * flags: bit0 tree-prior term (c[1]), bit1 data term (c[2]), bit2 condition_on_survival,
 * bit3 fp32 pruning kernels.
 * ll_tree / ll_data [U] may be NULL.  Eight kernel launches, no host synchronisation.            */
#define VBSKY_TERM_TREE 1
#define VBSKY_TERM_DATA 2
#define VBSKY_CONDITION_ON_SURVIVAL 4
#define VBSKY_PRUNE_F32 8 /* evaluate the data term with the fp32 pruning kernels */
size_t vbsky_loglik_workspace_bytes(const vbsky_layout* lay, const vbsky_tips* tips, int32_t U, int32_t m);
int vbsky_loglik_value_and_grad(const vbsky_layout* lay, const vbsky_tips* tips, int32_t U,
                                const int32_t* unit_tree, int32_t m, const vbsky_params* in,
                                const vbsky_subst_model* model, int32_t flags, double* ll,
                                double* ll_tree, double* ll_data, vbsky_param_grads* grads,
                                void* ws, size_t ws_bytes, void* stream);

/* Site-pattern-block sharding (north star: "the ensemble shards across trees and site-pattern blocks"; no counterpart
 * in the reference).  When there are too few (tree, sample) units to fill the GPUs (one big tree, few samples), the ranks
 * of a group evaluate the SAME units over disjoint ranges of 256-pattern tiles:
 *   vbsky_tips_tiles                number of tiles per tree (ceil(S / 256));
 *   vbsky_prune_value_and_grad_tiles  vbsky_prune_value_and_grad restricted to tiles [tile_begin, tile_end): ll and dll_dbl
 *                                   are PARTIAL sums over those patterns (site_ll is written for them only);
 *   vbsky_loglik_forward_partial    first half of vbsky_loglik_value_and_grad: height chain, tree prior (+ gradient),
 *                                   data term over the tile range -> partial [ll_data (U) | d/d bl (U x (2n-2))] in ws;
 *   vbsky_loglik_exchange_buffer    device pointer / length (doubles) of that contiguous region: sum it over the group
 *                                   (one all-reduce, ~U * (2n-1) doubles) between the two halves;
 *   vbsky_loglik_backward_complete  second half: combines the terms and runs the adjoint of the height chain.
 * vbsky_loglik_value_and_grad is exactly forward_partial(all tiles) + backward_complete.                          */
int32_t vbsky_tips_tiles(const vbsky_tips* tips);
int vbsky_prune_value_and_grad_tiles(const vbsky_layout* lay, const vbsky_tips* tips, int32_t U,
                                     const int32_t* unit_tree, const double* bl,
                                     const vbsky_subst_model* model, int32_t tile_begin,
                                     int32_t tile_end, double* site_ll, double* ll, double* dll_dbl,
                                     void* ws, size_t ws_bytes, void* stream);
int vbsky_loglik_forward_partial(const vbsky_layout* lay, const vbsky_tips* tips, int32_t U,
                                 const int32_t* unit_tree, int32_t m, const vbsky_params* in,
                                 const vbsky_subst_model* model, int32_t flags, int32_t tile_begin,
                                 int32_t tile_end, int32_t want_grad, vbsky_param_grads* grads,
                                 void* ws, size_t ws_bytes, void* stream);
double* vbsky_loglik_exchange_buffer(const vbsky_layout* lay, const vbsky_tips* tips, int32_t U,
                                     int32_t m, void* ws, int64_t* count);
int vbsky_loglik_backward_complete(const vbsky_layout* lay, const vbsky_tips* tips, int32_t U,
                                   const int32_t* unit_tree, int32_t m, const vbsky_params* in,
                                   const vbsky_subst_model* model, int32_t flags, double* ll,
                                   double* ll_tree, double* ll_data, vbsky_param_grads* grads,
                                   void* ws, size_t ws_bytes, void* stream);

/* The reduction step of one ensemble ELBO evaluation (no counterpart in the reference, which visits
 * trees one at a time, fasta.py:456-487): from the per-unit outputs of vbsky_loglik_value_and_grad build the payload
 * of the single all-reduce,
 *   out[0]                      = scale * sum_u ll[u]
 *   out[1], out[2]              = scale * sum_u d/d origin, d/d clock_rate
 *   out[3 + i*m + j]            = scale * sum_u d/d {R, delta, s, rho}[u][j]                  (i = 0..3; NULL => 0)
 *   out[3 + 4m + t*(n-1) + j]   = scale * sum_{u of tree t} d/d proportions[u][j] (j < n-2), d/d root_proportion[u] (j = n-2)
 * (scale = 1/K for the mean over Monte-Carlo draws, optim.py:74).  unit_tree [U] must be non-decreasing (units
 * grouped by tree).  Fixed summation order.  vbsky_reduce_elbo_count = 3 + 4m + T(n-1) doubles.             */
int64_t vbsky_reduce_elbo_count(int32_t T, int32_t n, int32_t m);
int vbsky_reduce_elbo(int32_t U, int32_t T, int32_t n, int32_t m, const int32_t* unit_tree,
                      const double* ll, const vbsky_param_grads* grads, double scale, double* out,
                      void* stream);

/* ------------------------------------------------------------------------------------------
 * Local variational family (the per-tree flows of fasta.py:378-384): reparameterised samples
 *   x = clip(expit(mu + exp(log_sigma) * eps), 1e-7, 1-1e-7)
 * (Transform(dim, Compose(DiagonalAffine, ZeroOne)) over MeanField: prob/transform.py:130-202,
 * prob/distribution.py:52-65) and log q(x) as TransformedDistribution.log_pdf computes it
 * (transform.py:56-63), for the entropy term of optim.loss (optim.py:54-65).
 * mu, log_sigma dev [T][d] per tree; eps dev [U][d] standard-normal base draws (supplied by the
 * caller); x dev [U][d]; logq dev [U].  backward: given dx [U][d] and dlogq [U] (either may be NULL)
 * writes the per-unit adjoints dmu, dlog_sigma [U][d] (the caller sums the units of a tree).     */
int vbsky_local_flow_sample(int32_t U, int32_t d, const int32_t* unit_tree, const double* mu,
                            const double* log_sigma, const double* eps, double* x, double* logq,
                            void* stream);
int vbsky_local_flow_backward(int32_t U, int32_t d, const int32_t* unit_tree, const double* mu,
                              const double* log_sigma, const double* eps, const double* dx,
                              const double* dlogq, double* dmu, double* dlog_sigma, void* stream);

/* One optimiser step of SeqData.loop (fasta.py:388, 421-433) fused over a flat block of n
 * variational parameters on the device: jax.example_libraries.optimizers.adagrad(step_size,
 * momentum = 0.9 by default) on nan_to_num(g):  g_sq += g^2;
 * m = (1 - momentum) g / sqrt(g_sq) + momentum m (0 where g_sq == 0);  x -= step_size m.       */
int vbsky_adagrad_update(int64_t n, double* x, const double* g, double* g_sq, double* m,
                         double step_size, double momentum, void* stream);

/* CUDA-event timing of the fused pruning kernel alone (used for the roofline figure): after
 * vbsky_prune_timing(1, NULL, NULL) every prune call of this thread brackets the kernel with an
 * event pair on its stream; a later call drains them into *ms_sum / *launches (and sets the new
 * enable state).                                                                               */
int vbsky_prune_timing(int32_t enable, double* ms_sum, int64_t* launches);

/* ------------------------------------------------------------------------------------------
 * Host-buffer entry point: what a caller that holds NumPy arrays (the reference's jitted
 * `step` receives its arguments that way, fasta.py:461-471) binds.  A session owns the device
 * staging buffers and workspace for up to U_max units.  The call copies unit_tree/bl host ->
 * device, runs vbsky_prune_value_and_grad, copies ll (and dll_dbl unless NULL) back and
 * synchronises the stream before returning.  Pinned host memory makes the copies asynchronous
 * DMA; pageable memory works too.                                                              */
typedef struct vbsky_session vbsky_session;
int vbsky_session_create(const vbsky_layout* lay, const vbsky_tips* tips, int32_t U_max,
                         int32_t m_max, vbsky_session** out);
void vbsky_session_destroy(vbsky_session* s);
/* vbsky_prune_value_and_grad with host buffers (unit_tree, bl in; ll, dll_dbl out). */
int vbsky_prune_value_and_grad_host(vbsky_session* s, int32_t U, const int32_t* unit_tree,
                                    const double* bl, const vbsky_subst_model* model, double* ll,
                                    double* dll_dbl, void* stream);
/* vbsky_loglik_value_and_grad with host buffers: every pointer in `in`, `grads`, ll, ll_tree,
 * ll_data and unit_tree is HOST memory here.                                                  */
int vbsky_loglik_value_and_grad_host(vbsky_session* s, int32_t U, const int32_t* unit_tree,
                                     int32_t m, const vbsky_params* in,
                                     const vbsky_subst_model* model, int32_t flags, double* ll,
                                     double* ll_tree, double* ll_data, vbsky_param_grads* grads,
                                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VBSKY_B200_H */
