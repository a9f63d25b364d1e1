/*
 * zkp_b200 -- C ABI of the B200-native zkplonk hot path.
 *
 * This is the drop-in boundary: the entry points a Rust `extern "C"` block in
 * the reference's `poly-commit` / `zksnarks` crates would bind to replace
 *   - poly_commit::Fft::{new,dft,idft,coset_dft,coset_idft}   (call sites src/prover.rs:87-88,
 *     121-124,192,229; src/prover/quotient_poly.rs:50-58,115,145,237; src/key.rs:83,121-131,
 *     222-245; src/permutation.rs:194-197,229-235)
 *   - PlonkParams::commit / poly_commit::msm_curve_addition    (src/prover.rs:133-136,194,
 *     262-265,440,452; src/key.rs:138-159; src/prover/proof.rs:507-526)
 *   - the element-wise prover rounds between them              (src/prover.rs:107-452,
 *     src/prover/quotient_poly.rs, src/permutation.rs:205-300,
 *     src/prover/linearization_poly.rs)
 * INTEGRATION.md shows the Rust-side binding.
 *
 * Layouts (explicit, because the Rust struct layout of the absent crates is not visible):
 *   Fr  : 4 x uint64 little-endian limbs, Montgomery form, R = 2^256  (src/lib.rs:583-588)
 *   Fq  : 6 x uint64 little-endian limbs, Montgomery form, R = 2^384
 *   G1 affine : x then y = 12 x uint64; the point at infinity is x = y = 0
 *
 * All functions return 0 on success or a negative ZKP_ERR_* code; nothing unwinds across
 * the boundary (the reference builds with panic = "abort", Cargo.toml:52).  A context owns
 * one CUDA device and one stream; calls on the same context are serialised by the caller
 * (one context per proving thread mirrors `Prover: Clone`, src/prover.rs:28).
 * There is NO CPU fallback: every entry point fails with ZKP_ERR_CUDA when no sm_100 device
 * is usable.
 */
#ifndef ZKP_B200_H
#define ZKP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZKP_OK 0
#define ZKP_ERR_INVALID (-1)   /* bad argument (null, size, k out of range)                      */
#define ZKP_ERR_CUDA (-2)      /* CUDA runtime failure; see zkp_last_error()                     */
#define ZKP_ERR_DEGREE (-3)    /* polynomial degree exceeds the SRS: PlonkParams::commit's Err   */
#define ZKP_ERR_NOMEM (-4)
#define ZKP_ERR_VERIFY (-6)    /* proof rejected (Error::ProofVerificationError / PairingCheckFailure)   */
#define ZKP_ERR_STATE (-5)     /* call out of order (zkp_prover_prove_witness before set_wiring)   */

typedef struct zkp_ctx zkp_ctx;
typedef struct zkp_srs zkp_srs;   /* device-resident [tau^i]_1 powers (PlonkParams after trim)  */
typedef struct zkp_buf zkp_buf;   /* device-resident vector of Fr                               */
typedef struct zkp_comm zkp_comm; /* this rank's link to the GPUs sharing one job (NCCL)         */

/* ---- context --------------------------------------------------------------------- */
int zkp_ctx_create(int device, zkp_ctx** out);
void zkp_ctx_destroy(zkp_ctx* ctx);
const char* zkp_last_error(const zkp_ctx* ctx); /* text of the last CUDA failure          */
const char* zkp_strerror(int code);
int zkp_ctx_sync(zkp_ctx* ctx);
void* zkp_ctx_stream(zkp_ctx* ctx);             /* cudaStream_t the kernels are launched on */
int zkp_sm_count(const zkp_ctx* ctx);
/* number of kernels this library launched on ctx since creation (bench.py gpu_launches) */
uint64_t zkp_launch_count(const zkp_ctx* ctx);
/* points summed by the MSM since creation (algorithmic-work accounting of the roofline report) */
uint64_t zkp_msm_point_count(const zkp_ctx* ctx);

/* CUDA-event timing on the context's stream (events are recorded where the kernels run) */
int zkp_timer_start(zkp_ctx* ctx);
int zkp_timer_stop_ms(zkp_ctx* ctx, float* ms);  /* records stop, synchronises, returns ms  */

/* Per-kernel timing for the roofline report: when enabled, the library brackets its named
 * kernel groups ("msm_accumulate", "msm_sort", "msm_reduce", "ntt_pass", ...) with CUDA
 * events on the context's stream.  zkp_prof_read synchronises and sums a group. */
int zkp_prof_enable(zkp_ctx* ctx, int on);
int zkp_prof_reset(zkp_ctx* ctx);
int zkp_prof_read(zkp_ctx* ctx, const char* name, float* total_ms, uint64_t* count);

/* ---- Fr vectors on the device ----------------------------------------------------- */
int zkp_buf_alloc(zkp_ctx* ctx, size_t n, zkp_buf** out);
int zkp_buf_free(zkp_ctx* ctx, zkp_buf* buf);
size_t zkp_buf_len(const zkp_buf* buf);
int zkp_buf_upload(zkp_ctx* ctx, zkp_buf* dst, size_t dst_off, const uint64_t* src, size_t n);
int zkp_buf_download(zkp_ctx* ctx, const zkp_buf* src, size_t src_off, uint64_t* dst, size_t n);
/* Strided copies between a host matrix with a row pitch of host_pitch Fr and a dense device matrix of height x
 * width Fr: a rank's column slab of a natural-order host vector and back (four-step NTT). */
int zkp_buf_upload_2d(zkp_ctx* ctx, zkp_buf* dst, size_t dst_off, const uint64_t* src, size_t width, size_t height,
                      size_t host_pitch);
int zkp_buf_download_2d(zkp_ctx* ctx, const zkp_buf* src, size_t src_off, uint64_t* dst, size_t width, size_t height,
                        size_t host_pitch);
/* Wrap device memory owned by the caller (e.g. the tensor an NCCL collective reads and writes) as
 * a zkp_buf; zkp_buf_free on it releases only the handle. */
int zkp_buf_wrap(zkp_ctx* ctx, void* device_ptr, size_t n, zkp_buf** out);
int zkp_buf_zero(zkp_ctx* ctx, zkp_buf* buf, size_t off, size_t n);
int zkp_buf_copy(zkp_ctx* ctx, zkp_buf* dst, size_t dst_off, const zkp_buf* src, size_t src_off,
                 size_t n);

/* Keccak-f[1600] permutation (host code) for the Merlin transcript the prover keeps on the host. */
void zkp_keccak_f1600(uint64_t state[25]);
/* Page-locked host memory for vectors that cross PCIe (the host-buffer entry points copy at
 * full link rate only from pinned memory). */
int zkp_host_alloc(size_t bytes, void** out);
int zkp_host_free(void* p);

/* ---- NTT: poly_commit::Fft ---------------------------------------------------------- */
/* Host vector, in place.  data holds 2^k Fr; the first len_in are inputs, the rest is
 * treated as zero (Fft pads short inputs).  inverse=0,coset=0: dft; 1,0: idft (x n^-1);
 * 0,1: coset_dft (a_i * g^i first, g = 7); 1,1: coset_idft (x g^-i last).
 * Output in natural order: out[j] = sum_i a_i w^(ij), w = Fft::new(k).generator(). */
int zkp_ntt(zkp_ctx* ctx, uint64_t* data, size_t len_in, unsigned k, int inverse, int coset);
/* Device-resident variant.  in may equal out.  in holds >= len_in, out >= 2^k elements. */
int zkp_ntt_dev(zkp_ctx* ctx, const zkp_buf* in, size_t len_in, zkp_buf* out, unsigned k,
                int inverse, int coset);
/* `batch` transforms of the same shape: polynomial b reads in[b*in_stride ..] and writes
 * out[b*out_stride ..] (wire / selector batches of src/prover.rs:121-124, src/key.rs:121-131). */
int zkp_ntt_dev_batch(zkp_ctx* ctx, const zkp_buf* in, size_t in_stride, size_t len_in,
                      zkp_buf* out, size_t out_stride, unsigned k, int inverse, int coset,
                      unsigned batch);
/* Building blocks of the multi-GPU four-step NTT (sharding.py FourStepNtt):
 * zkp_permute_dev   out[a][b][0..w) = in[b][a][0..w)   (B x A matrix of w-element blocks, not in place)
 * zkp_scale_matrix_dev  data[a][b] *= base1^((a0+a) b)            (mode 0: four-step twiddle)
 *                       data[a][b] *= base1^(a0+a) * base2^b      (mode 1: coset scaling) */
int zkp_permute_dev(zkp_ctx* ctx, const zkp_buf* in, size_t in_off, zkp_buf* out, size_t out_off, size_t A, size_t B,
                    size_t w);
int zkp_scale_matrix_dev(zkp_ctx* ctx, zkp_buf* data, size_t off, size_t rows, size_t cols, size_t a0,
                         const uint64_t base1[4], const uint64_t base2[4], int mode);
/* Fft accessors: generator w (kind 0), w^-1 (1), n^-1 (2), coset g (3), g^-1 (4) */
int zkp_fft_constant(unsigned k, int kind, uint64_t out[4]);
/* Fft::elements: out[i] = w^i, i < 2^k (device buffer) */
int zkp_fft_elements_dev(zkp_ctx* ctx, unsigned k, zkp_buf* out);

/* ---- KZG10 commit: PlonkParams ------------------------------------------------------- */
/* Upload n affine powers (n x 12 uint64).  Mirrors PlonkParams::trim: the handle is what
 * `keypair` holds afterwards.  The device expands them once into the window table
 * 2^(c w) * P_i (W = floor(255/c) + 1 rows, 96 * W * n bytes of HBM) that every later commit uses. */
int zkp_srs_load(zkp_ctx* ctx, const uint64_t* xy, size_t n, zkp_srs** out);
int zkp_srs_free(zkp_ctx* ctx, zkp_srs* srs);
size_t zkp_srs_len(const zkp_srs* srs);
/* Synthetic SRS [tau^i * G]_{i<n} generated on the device (PlonkParams::setup's structure,
 * tests/range.rs:26; tau is Montgomery Fr).  Used by benchmarks and full-size tests. */
int zkp_srs_generate(zkp_ctx* ctx, const uint64_t tau[4], size_t n, zkp_srs** out);
/* Powers [tau^(first + i) * G]_{i<n}: one rank's slice of a sharded SRS. */
int zkp_srs_generate_range(zkp_ctx* ctx, const uint64_t tau[4], size_t first, size_t n, zkp_srs** out);
/* PlonkParams::trim (src/key.rs:82; tests/range.rs:24-26) on the device: the first `keep` powers as a new SRS
 * (window-table rows sliced device-to-device when the window width is unchanged, else rebuilt). */
int zkp_srs_trim(zkp_ctx* ctx, const zkp_srs* srs, size_t keep, zkp_srs** out);
int zkp_srs_download(zkp_ctx* ctx, const zkp_srs* srs, size_t off, uint64_t* xy, size_t n);

/* msm_curve_addition(&bases[..n], &scalars[..n]) -> affine (Commitment::new).
 * scalars: n x 4 uint64 Montgomery Fr on the host. */
int zkp_msm_g1(zkp_ctx* ctx, const zkp_srs* srs, const uint64_t* scalars, size_t n,
               uint64_t out_xy[12]);
/* Same with device-resident scalars scalars[off .. off+n). */
int zkp_msm_g1_dev(zkp_ctx* ctx, const zkp_srs* srs, const zkp_buf* scalars, size_t off,
                   size_t n, uint64_t out_xy[12]);
/* PlonkParams::commit(&Coefficients): as zkp_msm_g1, but trailing zero coefficients are
 * ignored and ZKP_ERR_DEGREE is returned when the highest non-zero index is >= the SRS
 * length (the only way create_proof fails for an unsatisfied circuit).  An
 * all-zero polynomial commits to the identity (x = y = 0). */
int zkp_commit(zkp_ctx* ctx, const zkp_srs* srs, const uint64_t* coeffs, size_t n,
               uint64_t out_xy[12]);
int zkp_commit_dev(zkp_ctx* ctx, const zkp_srs* srs, const zkp_buf* coeffs, size_t off, size_t n,
                   uint64_t out_xy[12]);
/* Up to 8 commitments against the same SRS in one set of launches (the four wire commits of round
 * 1, the four quotient chunks of round 3: src/prover.rs:133-136,262-265).  out_xy: count x 12;
 * status[i] = ZKP_OK or ZKP_ERR_DEGREE per polynomial. */
struct zkp_poly_ref;
int zkp_commit_batch_dev(zkp_ctx* ctx, const zkp_srs* srs, const struct zkp_poly_ref* polys, unsigned count,
                         uint64_t* out_xy, int* status);
/* Index of the highest non-zero coefficient of coeffs[off .. off+n), -1 for the zero polynomial
 * (Coefficients::degree; what commit's degree check looks at). */
int zkp_poly_degree_dev(zkp_ctx* ctx, const zkp_buf* coeffs, size_t off, size_t n, long long* top);
/* MSM tuning knob (window bits c; 0 = automatic).  The window structure is baked into the SRS
 * table when it is loaded, so this applies to SRS handles created afterwards. */
int zkp_msm_set_window(zkp_ctx* ctx, unsigned c);

/* ---- prover rounds on device-resident polynomials ------------------------------------ */
/* A polynomial / evaluation vector living inside a device buffer: buf[off .. off+len). */
typedef struct zkp_poly_ref {
    const zkp_buf* buf;
    size_t off;
    size_t len;
} zkp_poly_ref;

/* zkp_ntt_dev on a sub-range: reads in.len inputs (zero-padded to 2^k), writes 2^k outputs at
 * out[out_off ..).  In place when both name the same storage. */
int zkp_ntt_ref_dev(zkp_ctx* ctx, zkp_poly_ref in, zkp_buf* out, size_t out_off, unsigned k, int inverse,
                    int coset);
int zkp_buf_fill(zkp_ctx* ctx, zkp_buf* buf, size_t off, size_t n, const uint64_t value[4]);
/* Coefficients::blind(h, rng) with the h+1 = count (<= 3) scalars drawn by the caller
 * (src/prover.rs:126-129,193): p <- p + (b0 + b1 X + ..)(X^n - 1); buf needs n + count slots. */
int zkp_poly_blind_dev(zkp_ctx* ctx, zkp_buf* buf, size_t off, size_t n, const uint64_t* blinders,
                       unsigned count);
/* compute_permutation_lagrange (src/permutation.rs:140-169): out[i] = K[wire] * w^gate with
 * enc[i] = wire << 30 | gate (host array), K = (1, 7, 13, 17), roots = Fft::elements. */
int zkp_perm_lagrange_dev(zkp_ctx* ctx, unsigned k, const uint32_t* enc, size_t n, const zkp_buf* roots,
                          zkp_buf* out, size_t out_off);
/* Permutation::compute_permutation_vec (src/permutation.rs:205-300): z[0] = 1,
 * z[i+1] = z[i] * prod_j (w_j[i] + beta K_j w^i + gamma) / prod_j (w_j[i] + beta sigma_j[i] + gamma).
 * wires / sigmas: evaluations over the n-point domain. */
int zkp_perm_z_dev(zkp_ctx* ctx, size_t n, const zkp_poly_ref wires[4], const zkp_poly_ref sigmas[4],
                   const zkp_buf* roots, const uint64_t beta[4], const uint64_t gamma[4], zkp_buf* out,
                   size_t out_off);

/* quotient_poly::compute between its NTTs (src/prover/quotient_poly.rs:74-114): all inputs are
 * evaluations on the 8n coset; out[i] = (gates_i + PI_i + permutation_i) / Z_H(g w8^i). */
typedef struct zkp_quotient_args {
    zkp_poly_ref wires[4];        /* a, b, c, d */
    zkp_poly_ref z, pi, l1;       /* z, public inputs, L1 * alpha^2 */
    zkp_poly_ref sel[11];         /* q_m q_l q_r q_o q_c q_4 q_arith q_range q_logic q_fixed q_var */
    zkp_poly_ref sigma[4];
    zkp_poly_ref linear;          /* the polynomial X (permutation.linear_evaluations) */
    uint64_t challenges[7][4];    /* alpha beta gamma range logic fixed-base var-base separation */
    uint64_t zh_inv[8][4];        /* 1 / Z_H on the coset: period 8 */
    uint32_t widget_mask;         /* bit0 range, 1 logic, 2 fixed-base, 3 var-base: clear = selector
                                     polynomial identically zero, widget skipped (contributes 0) */
    uint32_t coset_log_n;         /* non-zero (= log2 n): COSET layout.  Every vector holds the evaluations on a run
                                     of whole cosets (g w_8n^u) H_n of the 8n domain, u = coset_first .. , n each:
                                     element u_local * n + m is the point g w_8n^u w_n^m.  "Next gate" is m + 1 inside
                                     the coset and 1 / Z_H is constant on it, so a rank that owns cosets evaluates
                                     its part of the quotient with no data from any other rank; k8 is ignored and
                                     zkp_quotient_range_dev's count is the number of local points */
    uint32_t coset_first;
    uint32_t sliced;              /* zkp_quotient_range_dev only: wires, z, pi, l1 hold just the evaluations
                                     [first, first + count + 8) (indices mod 8n: the 8-element halo is the
                                     "next gate" of the last points) -- what a rank receives when the coset
                                     transforms are dealt out over GPUs and exchanged by all-to-all */
} zkp_quotient_args;
int zkp_quotient_dev(zkp_ctx* ctx, unsigned k8, const zkp_quotient_args* args, zkp_buf* out, size_t out_off);
/* Only the evaluations [first, first + count) of the same vector (out[i] for those i): one rank's
 * slice when the 8n points are split over GPUs; inputs are full-length on every rank. */
int zkp_quotient_range_dev(zkp_ctx* ctx, unsigned k8, const zkp_quotient_args* args, size_t first, size_t count,
                           zkp_buf* out, size_t out_off);

/* Coefficients::evaluate for up to 32 polynomials at one point (linearization_poly.rs:52-73). */
int zkp_poly_eval_dev(zkp_ctx* ctx, const zkp_poly_ref* polys, unsigned count, const uint64_t point[4],
                      uint64_t* out /* count x 4 */);
/* The same for up to 32 polynomials at up to two points in one launch and one read-back: polynomial i
 * is evaluated at points[which[i]] (which = NULL: all at points[0]).  The round driver opens every
 * polynomial of a proof -- at z and at z w -- with one call. */
int zkp_poly_eval2_dev(zkp_ctx* ctx, const zkp_poly_ref* polys, const uint8_t* which, unsigned count,
                       const uint64_t points[8], uint64_t* out /* count x 4 */);
/* out[i] = sum_k scalars[k] * polys[k][i], i < out_len (polys zero beyond their len; count <= 16):
 * widget.linearize sums, t_low + z^n t_mid + .., sum_i v^i p_i. */
int zkp_poly_lincomb_dev(zkp_ctx* ctx, const zkp_poly_ref* polys, const uint64_t* scalars, unsigned count,
                         zkp_buf* out, size_t out_off, size_t out_len);
/* ruffini: quotient of in(X) by (X - point), remainder dropped; writes in.len - 1 coefficients
 * (PlonkParams::compute_aggregate_witness, src/prover.rs:422-451). */
int zkp_poly_div_linear_dev(zkp_ctx* ctx, zkp_poly_ref in, const uint64_t point[4], zkp_buf* out,
                            size_t out_off);

/* ---- Prover::create_proof as one call (src/prover.rs:67-474) --------------------------- */
/* ProvingKey (fields src/key.rs:247-302) as device-resident vectors.  Index order of poly / eval8:
 * q_m q_l q_r q_o q_c q_4 q_arith q_range q_logic q_fixed_group_add q_variable_group_add
 * s_sigma_1 s_sigma_2 s_sigma_3 s_sigma_4.  poly[i]: n coefficients; eval8[i]: 8n coset evaluations. */
typedef struct zkp_proving_key {
    unsigned k;                    /* n = 2^k gates (padded) */
    zkp_poly_ref poly[15];
    zkp_poly_ref eval8[15];
    zkp_poly_ref linear8;          /* the polynomial X on the 8n coset */
    zkp_poly_ref sigma_evals[4];   /* sigma_j over the n-domain (round 2 reads them directly) */
    const zkp_buf* roots;          /* Fft::elements of the n-domain */
    uint64_t zh_inv[8][4];         /* 1 / Z_H on the coset (period 8) */
    uint64_t generator[4];         /* w of the n-domain (verifier_key.generator), Montgomery */
    uint32_t widget_mask;          /* as in zkp_quotient_args */
} zkp_proving_key;

typedef struct zkp_prover zkp_prover;   /* `Prover`: key + workspace + second stream */

/* The key's buffers and the SRS must outlive the prover.  Allocates the proof workspace
 * (about 70 n Fr) once; proofs on one prover are serialised by the caller. */
int zkp_prover_create(zkp_ctx* ctx, const zkp_srs* srs, const zkp_proving_key* key, zkp_prover** out);
int zkp_prover_destroy(zkp_prover* prover);
/* One proof.  transcript: the Merlin state after Transcript::base and the public-input appends
 * (src/prover.rs:54-55,99-105), serialised as 200 Keccak state bytes + pos, pos_begin, cur_flags.
 * Witness: the four wire columns over the n-domain, 4n x 4 uint64 Montgomery (a | b | o | d), from
 * host memory (pinned for full link rate) or already on the device; public inputs likewise as a
 * dense n-vector.  blinders: the 11 scalars `blind` draws, in draw order (src/prover.rs:126-129,193).
 * Out: the 11 commitments (a b c d z t_low t_mid t_high t_4 w_z w_zw; 12 uint64 each, affine
 * Montgomery, zeros = identity), the 16 evaluations in `Evaluations` order (Montgomery), and, if
 * non-null, the 1040-byte wire format (src/prover/proof.rs:36-66) and the transcript state after
 * the proof.  Returns ZKP_ERR_DEGREE where the reference returns Err (unsatisfied circuit). */
int zkp_prover_prove(zkp_prover* prover, const uint8_t transcript[203], const uint64_t* wires_host,
                     const zkp_buf* wires_dev, const uint64_t* pi_host, const zkp_buf* pi_dev,
                     const uint64_t blinders[44], uint64_t commitments[132], uint64_t evaluations[64],
                     uint8_t proof_bytes[1040], uint8_t transcript_out[203]);

/* ---- one job over several GPUs (BASELINE north_star) ------------------------------------
 * One rank (process or thread) per GPU, G in {1, 2, 4, 8}.  What is split:
 *   - every KZG commitment by SRS ranges: rank r multiplies coefficients [r L / G, (r + 1) L / G) with the
 *     matching powers; the G partial sums are all-gathered (192 bytes each) and added on every rank;
 *   - the 8n-point quotient domain by COSETS: rank r owns the 8 / G cosets (g w_8n^u) H_n, u in
 *     [8 r / G, 8 (r + 1) / G).  Each is an n-point coset transform of replicated coefficients (no exchange
 *     going in), the quotient is point-wise within a coset, and the inverse transform needs ONE exchange: slab s
 *     of every locally inverted coset goes to rank s, which combines the eight and ends up holding
 *     coefficients [s n / G, (s + 1) n / G) of every n-chunk of t(X) -- exactly the range of t_low / t_mid /
 *     t_high / t_4 its SRS range commits.
 * rank 0 makes the id, the host ships its 256 bytes to the other ranks, all ranks call zkp_comm_create
 * (collective).  nranks == 1 needs no id and no NCCL: the same code path on one GPU. */
int zkp_comm_unique_id(uint8_t out[256]);
int zkp_comm_create(zkp_ctx* ctx, const uint8_t id[256], int rank, int nranks, zkp_comm** out);
int zkp_comm_destroy(zkp_comm* comm);
int zkp_comm_rank(const zkp_comm* comm);
int zkp_comm_size(const zkp_comm* comm);
int zkp_comm_stats(const zkp_comm* comm, uint64_t* collectives, uint64_t* bytes_sent);
/* All-to-all of equal blocks of `count` Fr between device vectors (block p of src -> rank p; block r of dst <-
 * rank r), queued on the context's stream: the transpose step of the four-step NTT. */
int zkp_comm_all_to_all_dev(zkp_comm* comm, const zkp_buf* src, size_t src_off, zkp_buf* dst, size_t dst_off,
                            size_t count);
/* Four-step twiddle fused into the transpose: out[b][a] = in[a][b] * w_N^(+-(a0 + a) b) for an a-major rows x cols
 * matrix, N = 2^k (two-level twiddle table, one pass over the data). */
int zkp_twiddle_transpose_dev(zkp_ctx* ctx, const zkp_buf* in, size_t in_off, zkp_buf* out, size_t out_off, size_t rows,
                              size_t cols, size_t a0, unsigned k, int inverse);
/* zkp_commit_batch_dev over all ranks of `comm` (polynomials replicated, SRS complete on every rank):
 * collective; every rank returns the same commitments / status. */
int zkp_commit_batch_sharded_dev(zkp_ctx* ctx, zkp_comm* comm, const zkp_srs* srs, const zkp_poly_ref* polys,
                                 unsigned count, uint64_t* out_xy /* count x 12 */, int* status /* count */);
/* Fft::coset_dft restricted to whole cosets of the n-domain inside the 8n domain:
 * out[(u - first) n + m] = p(g w_8n^u w_n^m) for u = first .. first + count - 1, m < n = 2^k.
 * len_in <= 2n coefficients (a blinded polynomial has n + 3).  Concatenated over u = 0 .. 7 and read as
 * out[u][m] -> index 8 m + u this is the reference's 8n-point coset_dft. */
int zkp_coset8_ntt_dev(zkp_ctx* ctx, const zkp_buf* in, size_t in_off, size_t len_in, zkp_buf* out, size_t out_off,
                       unsigned k, unsigned first, unsigned count);
/* zkp_prover_create for rank zkp_comm_rank(comm) of a sharded proof: key->eval8[i] / linear8 hold this
 * rank's cosets only (8 n / G evaluations each, zkp_coset8_ntt_dev layout), everything else is as for one
 * GPU and replicated; the SRS is complete on every rank.  zkp_prover_prove / _prove_witness are then
 * collective: every rank calls them with the same inputs and returns the same proof. */
int zkp_prover_create_sharded(zkp_ctx* ctx, zkp_comm* comm, const zkp_srs* srs, const zkp_proving_key* key,
                              zkp_prover** out);

/* The same with the witness gather on the device (src/prover.rs:109-119, src/lib.rs:206-219): the
 * circuit's wiring is set once -- wire_idx[j * m + i] = witness index of wire j (a, b, o, d) at gate
 * i < m, pi_idx[c] = gate of public input c -- and each proof ships only the witness values
 * (num_w x 4 uint64 Montgomery) and the public-input values in the form the dense vector holds them. */
int zkp_prover_set_wiring(zkp_prover* prover, const uint32_t* wire_idx, size_t m, const uint32_t* pi_idx,
                          size_t pi_count);
int zkp_prover_prove_witness(zkp_prover* prover, const uint8_t transcript[203], const uint64_t* witness,
                             size_t num_w, const uint64_t* pi_values, const uint64_t blinders[44],
                             uint64_t commitments[132], uint64_t evaluations[64], uint8_t proof_bytes[1040],
                             uint8_t transcript_out[203]);

/* Host-side pieces of the driver, exported so they are testable without a GPU:
 * Merlin append_message / challenge_bytes on a serialised state, the scalar side of the
 * linearisation (challenges = alpha beta gamma range logic fixed var z; evals in `Evaluations`
 * order; out = scalars of q_m q_l q_r q_o q_4 q_c q_range q_logic q_fixed q_var z s_sigma_4),
 * compressed G1 and the 64-byte wide reduction of challenge_scalar. */
/* Transcript::new(label) (merlin): the 203-byte state IS the transcript object of a host that keeps
 * Merlin inside this library (the merlin crate does not expose its STROBE state). */
int zkp_transcript_init(uint8_t state[203], const uint8_t* label, uint32_t len);
int zkp_transcript_append(uint8_t state[203], const char* label, const uint8_t* msg, uint32_t len);
int zkp_transcript_challenge(uint8_t state[203], const char* label, uint8_t* out, uint32_t len);
int zkp_linearization_scalars(unsigned k, const uint64_t challenges[32], const uint64_t evals[60],
                              uint64_t out[48]);
int zkp_g1_compress(const uint64_t xy[12], uint8_t out[48]);
int zkp_fr_from_wide(const uint8_t bytes[64], uint64_t out_mont[4]);

/* ---- verifier glue (host code) ------------------------------------------------
 * VerificationKey (src/key.rs:203-214, fields of zksnarks::plonk::VerificationKey): the 15 commitments in the
 * order of zkp_proving_key.poly, the padded size 2^k and the constraint count the transcript was seeded with. */
typedef struct zkp_verifier_key {
    unsigned k;
    uint64_t constraints;
    uint64_t commitments[15][12];   /* affine Montgomery, zeros = identity */
} zkp_verifier_key;
/* [s]_2 / [s]_1 for a Montgomery-form scalar: the G2 half of the opening key PlonkParams::verification_key()
 * returns (EvaluationKey { g, h, beta_h }, src/commitment_scheme.rs:51-58): beta_h = [tau]_2 for an SRS made
 * from tau, h = [1]_2.  G2 affine as x0 x1 y0 y1 (6 uint64 each, Montgomery), zeros = identity. */
int zkp_g2_generator_mul(const uint64_t scalar[4], uint64_t out[24]);
/* Untrusted bytes -> validated values.  zkp_g1_decompress: 48-byte compressed G1 (zcash / dusk encoding) -> affine
 * Montgomery, ZKP_ERR_INVALID for a missing flag, x >= p, stray bits with the infinity flag, x off the curve or a
 * point outside the prime-order subgroup.  zkp_proof_decode: the 1040-byte proof (src/prover/proof.rs:36-66) ->
 * commitments / evaluations in the layout of zkp_prover_prove and zkp_verify, every point checked as above and
 * every scalar required to be canonical ("subgroup checks are done when the proof is deserialized", proof.rs:77). */
int zkp_g1_decompress(const uint8_t in[48], uint64_t out_xy[12]);
int zkp_proof_decode(const uint8_t bytes[1040], uint64_t commitments[132], uint64_t evaluations[64]);
int zkp_g1_generator_mul(const uint64_t scalar[4], uint64_t out[12]);
/* prod_i e(g1_i, g2_i) == 1: ZKP_OK or ZKP_ERR_VERIFY (multi_miller_loop(..).final_exp() == identity). */
int zkp_pairing_check(const uint64_t* g1 /* count x 12 */, const uint64_t* g2 /* count x 24 */, size_t count);
/* Self-check: for e(g1, g2) the Frobenius-based final exponentiation the checks use equals the cube of the plain
 * (p^12 - 1) / r square-and-multiply, and the pairing is not degenerate. */
int zkp_pairing_selftest(const uint64_t g1[12], const uint64_t g2[24]);
/* commitment_scheme::batch_check (src/commitment_scheme.rs:24-66): `count` opening proofs (point, commitment to the
 * witness, claimed evaluation, commitment to the polynomial) against the opening key; draws the "batch" challenge
 * from the transcript state (updated in place). */
int zkp_kzg_batch_check(const uint64_t beta_h[24], const uint64_t* points, const uint64_t* witness_comms,
                        const uint64_t* evals, const uint64_t* poly_comms, size_t count, uint8_t transcript[203]);
/* Verifier::verify (src/verifier.rs:46-81) + Proof::verify (src/prover/proof.rs:70-383).  transcript: the
 * verifier's base transcript (Transcript::base(label, vk, constraints)); the public inputs are appended here.
 * commitments / evaluations: as zkp_prover_prove returns them.  ZKP_OK = accepted, ZKP_ERR_VERIFY = rejected. */
int zkp_verify(const zkp_verifier_key* vk, const uint64_t beta_h[24], const uint8_t transcript[203],
               const uint64_t commitments[132], const uint64_t evaluations[64], const uint32_t* pi_idx,
               const uint64_t* pi_values, size_t pi_count);

#ifdef __cplusplus
}
#endif
#endif /* ZKP_B200_H */
