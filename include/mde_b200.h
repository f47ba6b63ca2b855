/*
 * mde_b200.h - C ABI of libmde_b200.so: the B200 (sm_100a) kernels behind the per-pixel depth
 * supervision / evaluation hot path of xeTaiz/mono-depth-estimation.
 *
 * Conventions (all entry points)
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary;
 *   - every buffer is CALLER-OWNED DEVICE memory unless the name says host; the library never
 *     allocates, never synchronises the stream and never throws;
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued on it and the call returns;
 *   - return value: MDE_OK (0) or a negative MDE_E* code; mde_last_error() gives the text
 *     (thread-local);
 *   - `ws` is a device workspace of mde_workspace_bytes(max_images) bytes, 16-byte aligned, that
 *     the caller initialises ONCE with mde_workspace_init(). Kernels leave it ready for the next
 *     call (self-cleaning accumulators), so no memset is needed between calls. One workspace
 *     must not be used by two streams at the same time, and n_img of any call must not exceed
 *     the max_images it was initialised with;
 *   - `pred_dtype` is MDE_F32 / MDE_F16 / MDE_BF16 (AMP: reference train.py:60,139 runs
 *     precision=16); arithmetic is always fp32 with fp64 cross-thread accumulation; gradients
 *     are written in the dtype of the tensor they belong to; targets are fp32;
 *   - tensors are contiguous NCHW; "n_img" = product of all leading dims of a depth map,
 *     "hw" = H*W pixels per image.
 *
 * Each entry point cites the reference interface it replaces (file:line in
 * /root/reference = xeTaiz/mono-depth-estimation).
 */
#ifndef MDE_B200_H
#define MDE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDE_OK 0
#define MDE_EINVAL (-1)   /* bad argument (null pointer, negative size, unknown enum) */
#define MDE_ECUDA (-2)    /* a CUDA runtime call failed; see mde_last_error() */
#define MDE_EALIGN (-3)   /* pointer not aligned to its element size */
#define MDE_ETOOBIG (-4)  /* size exceeds what the workspace / index type supports */

#define MDE_F32 0
#define MDE_F16 1
#define MDE_BF16 2

/* ---- metric suite ------------------------------------------------------------------------ */
/* raw per-image sums, one row of MDE_METRIC_NQ doubles per image */
enum {
  MDE_Q_NVALID = 0, /* #{t > 0} */
  MDE_Q_D1 = 1,     /* #{valid, max(p/t,t/p) < 1.25}      (IEEE fp32 divides, bit-exact) */
  MDE_Q_D2 = 2,     /* ... < 1.5625 */
  MDE_Q_D3 = 3,     /* ... < 1.953125 */
  MDE_Q_ABS = 4,    /* sum |p-t| */
  MDE_Q_SQ = 5,     /* sum (p-t)^2 */
  MDE_Q_LOG10 = 6,  /* sum |log10 p - log10 t| */
  MDE_Q_SLE = 7,    /* sum (log1p p - log1p t)^2 */
  MDE_Q_ABSREL = 8, /* sum |p-t|/t */
  MDE_Q_SQREL = 9,  /* sum (p-t)^2/t */
  MDE_Q_RSQ = 10,   /* sum sqrt((p-t)^2/t)      (reference's 'rmse' quirk, metrics.py:106-109) */
  MDE_Q_LNSQ = 11,  /* sum (ln p - ln t)^2 */
  MDE_METRIC_NQ = 12
};
/* finished metric values, fixed order */
enum {
  MDE_M_DELTA1 = 0, MDE_M_DELTA2 = 1, MDE_M_DELTA3 = 2, MDE_M_MAE = 3, MDE_M_MSE = 4,
  MDE_M_LOG10 = 5, MDE_M_MSLE = 6, MDE_M_ABSREL = 7, MDE_M_SQREL = 8, MDE_M_RMSE = 9,
  MDE_M_RMSE_TRUE = 10, /* sqrt(mse): not in the reference */
  MDE_M_RMSE_LOG = 11,  /* sqrt(sum (ln p-ln t)^2 / n): not in the reference */
  MDE_METRIC_NM = 12
};

/* flags for mde_metrics */
#define MDE_METRICS_REFERENCE_MATH 1u /* evaluate every metric with the reference's own op sequence
                                          (log10f/log1pf/IEEE divides everywhere); default uses the
                                          cheaper algebraically equal forms documented in DESIGN.md.
                                          The delta counts are bit-exact in both modes. */
/* which float-sum groups the caller needs (0 = all); sums of groups not requested come back 0 */
#define MDE_METRICS_NEED_LOG (1u << 8)   /* MDE_Q_LOG10, MDE_Q_LNSQ  -> log10, rmse_log */
#define MDE_METRICS_NEED_LOG1P (1u << 9) /* MDE_Q_SLE               -> msle */
#define MDE_METRICS_NEED_REL (1u << 10)  /* MDE_Q_ABSREL/SQREL/RSQ  -> absrel, sqrel, rmse */
#define MDE_METRICS_NEED_RSQ (1u << 11)  /* MDE_Q_RSQ alone         -> rmse (the reference's default
                                            train list, train.py:67, needs no absrel / sqrel) */

/*
 * Masked error metrics over a batch of depth maps in ONE pass (8 B/px).
 * Replaces MetricComputation.compute (reference metrics.py:58-67) and the metric functions
 * metrics.py:75-109,116-122: pred <- max(pred, 1e-7); valid = target > 0.
 *
 *   out_f64 layout (doubles):
 *     [0, NM)                 pooled values: one mean over ALL valid pixels of the call
 *                             (= what compute() returns for the call tensor)
 *     [NM, 2NM)               image-mean values: unweighted mean over images of per-image means
 *                             (= the reference's eval-loop semantics with batch size 1,
 *                             SURVEY 3.2); images without a valid pixel are skipped and counted
 *     [2NM, 2NM+NQ)           pooled raw sums
 *     [2NM+NQ]                number of images that had >= 1 valid pixel
 *     [2NM+NQ+1, 3NM+NQ+1)    sum over those images of the per-image values (image-mean x count), so that
 *                             out_f64[2NM ...] = {pooled raw sums, #images, per-image value sums} is ONE
 *                             contiguous vector a multi-GPU evaluation can all-reduce in place
 *   out_f32 (nullable): first 2NM entries of out_f64 rounded to fp32 (what callers log)
 *   per_image_values (nullable): [n_img][NM] doubles;  per_image_raw (nullable): [n_img][NQ]
 */
#define MDE_METRICS_OUT_F64 (3 * MDE_METRIC_NM + MDE_METRIC_NQ + 1)
int mde_metrics(const void* pred, int pred_dtype, const float* target, int64_t n_img, int64_t hw,
                unsigned flags, void* ws, double* out_f64, float* out_f32,
                double* per_image_values, double* per_image_raw, void* stream);

/*
 * Multi-GPU evaluation in ONE launch per rank (BASELINE configs[4]: 654 images sharded over the GPUs of a
 * box). The reference itself never combines metrics across ranks (metrics.py:19-39 logs without sync_dist; each rank of
 * pl.Trainer(gpus=N), train.py:137, averages its own shard); the evaluation result over the whole set - mean over images
 * of per-image means, metrics.py:35-41 + modules/base_module.py:71-76 - needs the ranks' {pooled raw sums, #valid
 * images, per-image value sums} added up: 25 doubles per rank. mde_metrics_sharded is mde_metrics whose finaliser does
 * that exchange itself over NVLink peer memory: it stores its 25 doubles into every peer's mailbox and adds the
 * world's rows of its own mailbox in rank order (bit-identical totals on every rank), so out_f64 / out_f32 come back
 * holding the values of the WHOLE sharded set; per_image_* stay local. No collective launch, no host involvement.
 *
 *   comm       nullptr: exactly mde_metrics. Otherwise a communicator of mde_peer_comm_create (a small DEVICE-resident
 *              descriptor: the world's mailboxes as mapped into this process, rank, world, timeout).
 *   mailboxes  MDE_PEER_MAILBOX_BYTES of device memory per rank, zeroed once (mde_peer_alloc), the own one included.
 *              Rows are tagged with `seq` and double-buffered by its parity.
 *   seq        0: the communicator numbers its calls itself (a counter in its device-resident descriptor, advanced by the
 *              finaliser - the launch can then be captured in a CUDA graph and replayed); otherwise the same on every
 *              rank for the same call and different from the previous call's (count up). Do not mix the two styles.
 *   n_img == 0 is legal here (a rank without images still owes the world its row); pred / target may then be null.
 *   A peer whose row does not arrive within the timeout (default 2000 ms) is given up on: NaN results and error flag 1
 *   in the workspace header instead of a hung GPU.
 * Every rank must make the matching call (same seq), as with any collective.
 */
#define MDE_MAX_PEERS 8
#define MDE_PEER_MAILBOX_BYTES 8192
#define MDE_PEER_HANDLE_BYTES 64
int mde_metrics_sharded(const void* pred, int pred_dtype, const float* target, int64_t n_img, int64_t hw,
                        unsigned flags, void* ws, double* out_f64, float* out_f32,
                        double* per_image_values, double* per_image_raw, void* comm, unsigned seq, void* stream);
/* Sum of n <= 31 doubles over the ranks through the same mailboxes (one warp): dst[i] = sum over ranks of src[i], in rank
 * order (bit-identical on every rank); zero_src != 0 clears src in the same launch (a rank's accumulator between two
 * exchanges, e.g. the pooled metric sums mde_loss_params.metrics_raw_accum collects - SURVEY 8e). src == dst is allowed
 * when zero_src == 0. seq as in mde_metrics_sharded; ws: the stream's workspace (receives the error flag on a timeout). */
int mde_peer_allreduce_f64(double* src, double* dst, int n, int zero_src, void* comm, unsigned seq, void* ws, void* stream);
int mde_peer_comm_create(void* const* mailboxes /* [world] */, int rank, int world, unsigned timeout_ms /* 0: 2000 */,
                         void** comm_out);
int mde_peer_comm_destroy(void* comm);
/* Mailbox plumbing (one process per GPU): allocate + zero a block on the current device (synchronous), export its
 * CUDA IPC handle (MDE_PEER_HANDLE_BYTES, to be handed to the other ranks by any means - torch.distributed's
 * all_gather_object in distributed.PeerComm), map a peer's block into this process (enables peer access lazily). */
int mde_peer_alloc(size_t bytes, void** ptr_out);
int mde_peer_free(void* ptr);
int mde_peer_export(void* ptr, unsigned char* handle_out);
int mde_peer_open(const unsigned char* handle, void** ptr_out);
int mde_peer_close(void* ptr);

/* Finish metric values from raw sums on the HOST (used after an all-reduce of raw sums across
 * ranks): values[NM] from raw[NQ]. Pure arithmetic, no CUDA. */
void mde_metrics_finalize_host(const double* raw, double* values);

/* ---- masked scalar losses, forward and backward fused -------------------------------------
 * Each call computes the loss (fp32, device scalar `loss_out`) and, if `grad` is non-null,
 * dloss/dpred * grad_scale in pred's dtype, in ONE cooperative launch: phase A reduces, a
 * grid-wide barrier publishes the totals, phase B writes the gradient (pred/target re-read
 * through L2). `totals_out` (nullable, doubles) receives the phase-A totals listed per loss.
 */
enum {
  MDE_LOSS_L1 = 0,          /* MaskedL1Loss      criteria.py:80-90   totals {n, sum|t-p|} */
  MDE_LOSS_MSE = 1,         /* MaskedMSELoss     criteria.py:67-77   totals {n, sum(t-p)^2} */
  MDE_LOSS_BERHU = 2,       /* berHuLoss         criteria.py:111-133 totals {n, sum a, n_hub, sum_hub a^2, c} */
  MDE_LOSS_LAINA_BERHU = 3, /* LainaBerHuLoss    criteria.py:476-506 totals {M, sum, dL/dc*M, n_argmax, c} */
  MDE_LOSS_SILOG = 4,       /* silog_loss        criteria.py:724-732 totals {n, sum d, sum d^2} */
  MDE_LOSS_EIGEN = 5,       /* MaskedDepthLoss   criteria.py:17-64   totals {D, A, my, ey, mx, ex} */
  MDE_LOSS_COUNT = 6
};
#define MDE_LOSS_NTOTALS 8

typedef struct {
  float variance_focus; /* silog: lambda (modules/bts.py:233 default 0.85) */
  float clamp_val;      /* laina: clamp_val (criteria.py:479, 1e-9) */
  int use_logs;         /* laina: use_logs (criteria.py:479) */
  int size_average;     /* laina: size_average (criteria.py:479) */
  float* metrics_accum; /* mde_masked_loss_metrics only, nullable: MDE_METRIC_NM device floats that receive
                           `+= pooled value` from the launch's finaliser - MetricComputation's running sums
                           (reference metrics.py:65-66) without a launch of their own */
  double* metrics_raw_accum; /* same, nullable: MDE_METRIC_NQ device doubles that receive `+= pooled raw sum`
                           (what a rank pools between two all-reduces of the metric sums) */
} mde_loss_params;

/*
 * kind      one of MDE_LOSS_*
 * pred      [n_img, h, w] in pred_dtype; target fp32 same shape
 * mask_u8   nullable; LainaBerHuLoss's optional third argument (criteria.py:485) as bytes
 * grad      nullable -> forward only
 * grad_scale  multiplies the gradient (autograd's grad_output when it is known up front; 1.0f otherwise)
 */
int mde_masked_loss(int kind, const void* pred, int pred_dtype, const float* target,
                    const uint8_t* mask_u8, int64_t n_img, int64_t h, int64_t w,
                    const mde_loss_params* params, float grad_scale, void* ws, float* loss_out,
                    double* totals_out, void* grad, void* stream);

/*
 * Same as mde_masked_loss (L1, MSE, BERHU, LAINA_BERHU, SILOG) with the POOLED metric suite of
 * mde_metrics fused into the reduce phase of the same launch: the training-step tail
 * `loss = criterion(pred, gt); loss.backward(); metric_logger.log_train(pred, gt, loss)`
 * (reference modules/eigen.py:31-33, modules/bts.py:106-108) costs 12 B/px instead of 12 + 8.
 * metric_flags: MDE_METRICS_NEED_* (fast metric forms only). metrics_f64 / metrics_f32 use the
 * layout of mde_metrics; only the pooled values [0,NM) and pooled raw sums [2NM, 2NM+NQ) are
 * formed (image-mean entries equal the pooled ones when n_img == 1 and are NaN otherwise).
 */
int mde_masked_loss_metrics(int kind, const void* pred, int pred_dtype, const float* target,
                            const uint8_t* mask_u8, int64_t n_img, int64_t h, int64_t w,
                            const mde_loss_params* params, float grad_scale, unsigned metric_flags,
                            void* ws, float* loss_out, double* totals_out, void* grad,
                            double* metrics_f64, float* metrics_f32, void* stream);

/*
 * Split-phase form of the same losses (L1, MSE, BERHU, LAINA_BERHU, SILOG) for GLOBAL-BATCH training over several
 * GPUs (the batch shards by image, the ranks exchange only the scalar totals, every rank ends with the
 * loss and the gradient of the single-GPU full-batch call - criteria.py:111-133 needs a max and then sums,
 * :724-732 the sums of d, d^2 and the count). Plain streaming launches; nothing waits on another rank:
 *
 *   partials = {0, 0, 0, 0, -inf, 0, 0, 0}                                       (device, 8 doubles, caller-owned)
 *   berHu / Laina:  mde_masked_loss_partials(kind, 0, ...)   partials[4] = max over the shard  -> all-reduce(MAX)
 *   all kinds:      mde_masked_loss_partials(kind, 1, ..., gmax_in = &partials[4], ...)
 *                                                            partials[0..3] += {S0, S1, N0, N1} -> all-reduce(SUM)
 *   mde_masked_loss_from_totals(kind, ..., totals = partials, grad_scale, loss_out, grad)
 *                                                            loss (same on every rank) + dloss/dpred of the shard
 * The totals are the ones mde_masked_loss reports through totals_out (same order).
 */
int mde_masked_loss_partials(int kind, int stage, const void* pred, int pred_dtype, const float* target,
                             const uint8_t* mask_u8, int64_t n, const mde_loss_params* params,
                             const double* gmax_in, double* partials, void* stream);
int mde_masked_loss_from_totals(int kind, const void* pred, int pred_dtype, const float* target,
                                const uint8_t* mask_u8, int64_t n, const mde_loss_params* params,
                                const double* totals, float grad_scale, float* loss_out, void* grad,
                                void* stream);

/* x[i] *= *scale_dev (device scalar) - applies a late-arriving grad_output to a stashed gradient */
int mde_scale_inplace(void* x, int dtype, int64_t n, const float* scale_dev, void* stream);

/* ---- DORN ordinal head -------------------------------------------------------------------- */
#define MDE_DISC_SID 0
#define MDE_DISC_UD 1

/*
 * OrdinalRegressionLayer.forward (reference network/Dorn.py:292-321).
 * x [n, 2K, hw] logits (pair k = channels 2k, 2k+1), clamp to [1e-8,1e4], P_k = softmax(a,b)[1];
 * decode[n,hw] int64 = #{k: P_k > 0.5} (bit-exact incl. clamp ties and 1-ulp near ties).
 * prob [n,K,hw] fp32 (nullable), decode (nullable).
 */
int mde_ordinal_layer_fwd(const void* x, int x_dtype, int64_t n, int64_t K, int64_t hw,
                          float* prob, int64_t* decode, void* stream);
/* backward of the layer: grad_x [n,2K,hw] (x's dtype) from grad_prob [n,K,hw] fp32 */
int mde_ordinal_layer_bwd(const void* x, int x_dtype, const float* grad_prob, int64_t n, int64_t K,
                          int64_t hw, void* grad_x, void* stream);

/* DORNModule.label_to_depth / depth_to_label (reference modules/dorn.py:95-107); label is
 * int64 (decode) for _i64 and fp32 otherwise. */
int mde_label_to_depth_i64(const int64_t* label, int64_t n, float alpha, float beta, int ord_num,
                           int discretization, float* depth, void* stream);
int mde_label_to_depth_f32(const float* label, int64_t n, float alpha, float beta, int ord_num,
                           int discretization, float* depth, void* stream);
int mde_depth_to_label(const float* depth, int64_t n, float alpha, float beta, int ord_num,
                       int discretization, float* label, void* stream);

/* ordLoss.forward (reference criteria.py:744-787): prob [n,K,hw] fp32, target [n,hw] fp32 SID
 * label; loss = -(sum_{k<=y} ln clamp(P) + sum_{k>y} ln clamp(1-P)) / (n*hw); grad_prob nullable. */
int mde_ord_loss(const float* prob, const float* target_label, int64_t n, int64_t K, int64_t hw,
                 float grad_scale, void* ws, float* loss_out, float* grad_prob, void* stream);

/*
 * Fused DORN supervision step: logits + metric gt depth -> decode, depth, loss and grad_logits
 * in one pass over the logits (1104 B/px at K=68). Equals, in the reference,
 *   decode, P = OrdinalRegressionLayer()(x)                 network/Dorn.py:292-321
 *   depth     = label_to_depth(decode)                      modules/dorn.py:95-100
 *   y_sid     = depth_to_label(gt)                          modules/dorn.py:102-107
 *   loss      = ordLoss()(P, y_sid); loss.backward()        criteria.py:744-787
 * Any of prob / decode / depth / grad_x may be null.
 */
int mde_dorn_fused(const void* x, int x_dtype, const float* gt_depth, int64_t n, int64_t K,
                   int64_t hw, float alpha, float beta, int discretization, float grad_scale,
                   void* ws, float* loss_out, float* prob, int64_t* decode, float* depth,
                   void* grad_x, void* stream);

/* OrdinalRegressionLoss.__call__ (reference criteria.py:789-836): prob [n,2K,hw] log-probabilities
 * laid out [K '<=' planes | K '>' planes]; label = trunc(K ln(gt/alpha)/ln(beta/alpha)); mean over
 * gt > 0 pixels. grad_prob nullable (same layout/dtype fp32). */
int mde_ordinal_regression_loss(const float* prob, const float* gt_depth, int64_t n, int64_t K,
                                int64_t hw, float alpha, float beta, int discretization,
                                float grad_scale, void* ws, float* loss_out, float* grad_prob,
                                void* stream);

/* ---- virtual-normal loss ------------------------------------------------------------------ */
/*
 * VNL_Loss.forward(gt_depth, pred_depth, select) (reference criteria.py:866-1045) with a
 * SUPPLIED triplet tensor (the select_index hook, criteria.py:912-932): trip [3, n_trip] int64
 * flat pixel indices (y*w + x), the same triplets for every image (criteria.py:948-950).
 * Back-projection, mask logic, the z==0 fix-up quirk (criteria.py:1004), per-triplet loss, the
 * exact 25 % trim (radix select on the fp32 bit pattern) and the scatter-add backward are fused
 * in one cooperative launch.
 *   scratch  device, mde_vnl_scratch_bytes(n_img, n_trip, h, w) bytes, 16-byte aligned (histograms, per-triplet losses and
 *            gradient factors, the depths re-laid pixel-major for the gathers; need not be zeroed)
 *   stats_out (nullable) doubles {M valid, q dropped, threshold, n_below, n_tie, kept_sum}
 */
size_t mde_vnl_scratch_bytes(int64_t n_img, int64_t n_trip, int64_t h, int64_t w);
int mde_vnl_loss(const float* gt_depth, const void* pred, int pred_dtype, const int64_t* trip,
                 int64_t n_img, int64_t h, int64_t w, int64_t n_trip, float fx, float fy, int select,
                 float grad_scale, void* ws, void* scratch, float* loss_out, double* stats_out,
                 void* grad, void* stream);

/*
 * Metrics of a prediction and a target that are BOTH bilinearly resized to out_h x out_w first, as the test steps of
 * the eigen / dorn / my modules do (F.interpolate(mode='bilinear'), align_corners=False: reference
 * modules/eigen.py:49-51, modules/dorn.py:181-183, modules/my.py:64-66) - sampled on the fly, the resized images are
 * never materialised. pred [n_img,pred_h,pred_w], target [n_img,target_h,target_w], fp32. Outputs as mde_metrics.
 */
int mde_metrics_resized(const float* pred, int64_t pred_h, int64_t pred_w, const float* target, int64_t target_h,
                        int64_t target_w, int64_t n_img, int64_t out_h, int64_t out_w, unsigned flags, void* ws,
                        double* out_f64, float* out_f32, double* per_image_values, double* per_image_raw,
                        void* stream);

/* ---- VNL's classification half ------------------------------------------ */
/*
 * WCEL_Loss.forward(pred_logit, gt_bins, gt) (reference criteria.py:839-863), forward + backward:
 *   loss = -sum_px sum_c weight[bin_px][c] * log_softmax(logits_px)[c] / #(gt > 0)
 * logits [n,C,hw] (x_dtype), gt_bins [n,hw] int32 (a bin outside [0,C) - modules/vnl.py:213 marks padding
 * C+1 - has an all-zero one-hot row), gt_depth [n,hw] fp32 (only its count of > 0 pixels is used),
 * weight [C,C] fp32 row-normalised as criteria.py:846-848 leaves it (row = gt bin), rowsum [C] = fp32 sums
 * of those rows. grad_logits (nullable, dtype of logits) = (softmax * rowsum[bin] - weight[bin]) / n_valid.
 * Two launches (valid count, fused pass); all pointers are device memory.
 */
int mde_wcel_loss(const void* logits, int x_dtype, const int* gt_bins, const float* gt_depth,
                  const float* weight, const float* rowsum, int64_t n, int64_t C, int64_t hw,
                  float grad_scale, void* ws, float* loss_out, void* grad_logits, void* stream);
/* VNLModule.depth_to_bins (reference modules/vnl.py:202-217): bins_out int32 [n]; depth_inout is clamped to
 * [depth_min, depth_max] IN PLACE and padding (depth < 0) restored to -1, as the reference mutates its input. */
int mde_depth_to_bins(float* depth_inout, int64_t n, float depth_min, float depth_max, float depth_min_log,
                      float depth_bin_interval, int64_t C, int* bins_out, void* stream);
/* VNLModule.bins_to_depth (reference modules/vnl.py:219-230): prob [n,C,hw] -> depth_out [n,hw] fp32 =
 * 10 ^ sum_c prob_c * border_c; and its backward grad_prob_c = grad_depth * ln(10) * depth * border_c. */
int mde_bins_to_depth(const void* prob, int x_dtype, const float* border, int64_t n, int64_t C, int64_t hw,
                      float* depth_out, void* stream);
int mde_bins_to_depth_bwd(const float* depth, const float* grad_depth, const float* border, int64_t n, int64_t C,
                          int64_t hw, int x_dtype, void* grad_prob, void* stream);

/* ---- MiDaS scale-and-shift alignment (evaluation side) ------------------- */
/*
 * compute_scale_and_shift(prediction, target, mask) (reference criteria.py:154-176): per image the least-squares
 * (scale, shift) that aligns pred to target over the mask (mask_u8 nullable: target > 0, :155-156); zeros where the
 * 2x2 system is singular (:170-174). pred [n_img,hw] (pred_dtype), target [n_img,hw] fp32, outputs fp32 [n_img].
 * The workspace must have been initialised for >= n_img images.
 */
int mde_scale_and_shift(const void* pred, int pred_dtype, const float* target, const uint8_t* mask_u8,
                        int64_t n_img, int64_t hw, void* ws, float* scale_out, float* shift_out,
                        double* sums_out /* nullable: [n_img][5] = a00, a01, a11, det, det != 0 */, void* stream);
/* MidasModule.scale_shift (reference modules/midas.py:56-62): out = scale[img] * pred + shift[img], fp32 out. */
int mde_apply_scale_shift(const void* pred, int pred_dtype, const float* scale, const float* shift, int64_t n_img,
                          int64_t hw, float* out, void* stream);

/*
 * MidasLoss.forward WITHOUT the scale/shift alignment (reference criteria.py:306-332 with loss in {'mse','l1','trim'},
 * reduction='batch-based'; the criterion of the registered method `my`, modules/my.py:39): data term
 * (mse_loss :219-223 or l1_loss :201-206; 'trim' :208-217 trims nothing as written and equals l1) plus
 * alpha * GradientLoss over `scales` strided grids (:226-244, :283-303), forward + backward in one cooperative
 * launch. data_kind: 0 mse, 1 l1/trim. grad nullable (dtype of pred). pred/target [n_img,h,w].
 * scale/shift (nullable, [n_img] fp32 from mde_scale_and_shift): the 'ssi' variants (criteria.py:326-328) evaluate the
 * loss on scale * pred + shift; grad is then dL/d(aligned pred) and mde_midas_ssi_backward finishes the chain rule.
 */
int mde_midas_loss(const void* pred, int pred_dtype, const float* target, const float* scale, const float* shift,
                   int64_t n_img, int64_t h, int64_t w, int data_kind, float alpha, int scales, float grad_scale,
                   void* ws, float* loss_out, void* grad, void* stream);
/* Backward of the 'ssi' variants through compute_scale_and_shift: grad_inout holds g = dL/d(aligned pred) (fp32) on
 * entry and dL/dpred on exit: s g_i + m_i (U y_i - 2 s U p_i - t U - s V), U, V from the per-image sums of
 * mde_scale_and_shift and G0 = sum g, G1 = sum g p. coef_scratch: device, n_img * 4 floats, 16-byte aligned. */
int mde_midas_ssi_backward(const float* pred, const float* target, const float* scale, const float* shift,
                           const double* sums, int64_t n_img, int64_t hw, void* ws, float* coef_scratch,
                           float* grad_inout, void* stream);

/*
 * TrimmedProcrustesLoss (reference criteria.py:335-363; `midas --loss ssitrim`, modules/midas.py:36-37), three calls:
 *
 * mde_robust_normalize = normalize_prediction_robust (criteria.py:135-152) of prediction AND target under the mask
 * target > 0: x' = (x - m) / s with m the lower median of mask * x over ALL pixels of the image (exact: radix select
 * on the fp32 bit pattern) and s = clamp(sum mask |x - m| / sum mask, 1e-6); m = 0, s = 1 for an image without a
 * valid pixel. stats_* [n_img][8] fp32 = {m, s, n_valid, k (bit pattern of the uint32 index that holds the median),
 * mask_k, Z = sum mask sign(x - m), s_was_clamped, 0}. pred/target/outputs fp32 [n_img, hw].
 */
size_t mde_robust_scratch_bytes(int64_t n_img);
int mde_robust_normalize(const float* pred, const float* target, int64_t n_img, int64_t hw,
                         void* scratch /* device, mde_robust_scratch_bytes(n_img), 8-byte aligned, any content */,
                         float* stats_pred, float* stats_target, float* pred_out, float* target_out, void* stream);
/* mde_midas_loss on tensors whose validity is NOT target > 0: valid_src > 0 decides (criteria.py:351 takes the mask
 * from the original target, the loss is evaluated on the normalised one). No scale/shift. Otherwise as mde_midas_loss. */
int mde_midas_loss_masked(const void* pred, int pred_dtype, const float* target, const float* valid_src,
                          int64_t n_img, int64_t h, int64_t w, int data_kind, float alpha, int scales,
                          float grad_scale, void* ws, float* loss_out, void* grad, void* stream);
/* Backward through the normalisation of the prediction: grad_inout holds g = dL/dx' on entry, dL/dx on exit:
 * g_j / s - mask_j sign(x'_j) Gx / (s n) + [j == k] mask_k (Gx Z / (s n) - G / s), G = sum g, Gx = sum g x' per image
 * (median gradient to element k, as torch.median(dim).values back-propagates; Gx terms dropped where s was clamped).
 * pred_norm: pred_out of mde_robust_normalize; target: the ORIGINAL target; coef_scratch: n_img * 4 floats, 16-byte aligned. */
int mde_robust_backward(const float* pred_norm, const float* target, const float* stats_pred, int64_t n_img,
                        int64_t hw, void* ws, float* coef_scratch, float* grad_inout, void* stream);

/* ---- depth map -> colours ------------------------------------------------ */
/*
 * colored_depthmap(depth, d_min, d_max, do_mapping) (reference visualize.py:8-17): rel = (depth - d_min) / (d_max -
 * d_min) * 255 in float32, cast to uint8 as numpy does, then OpenCV's COLORMAP_INFERNO (BGR). auto_range != 0: d_min /
 * d_max are the minimum / maximum of the map (NaN when it holds a NaN, as np.min / np.max), found on the device in
 * scratch4 (4 words, any content). out: uint8 [n, 3] (do_mapping) or [n], 4-byte aligned.
 */
int mde_colored_depthmap(const float* depth, int64_t n, float d_min, float d_max, int auto_range, int do_mapping,
                         unsigned* scratch4, uint8_t* out, void* stream);

/* ---- layered-depth ("stdepth") base criterion -------------------------- */
/*
 * The closure BaseModule.setup_criterion returns (reference modules/base_module.py:124-208; the criterion of the
 * registered methods `bts` and `laina`), forward + backward in one cooperative launch, for the terms that are masked
 * reductions: flags is a bit set of 1 depth_silog (:157/:160, depth_w * nan_to_num(silog_loss) over targ[D] > 0),
 * 2 color_mae (:158), 4 color_mse (:161), 8 all_mse (:163-164), 16 all_mae (:166-167), 32 fb_divergence (:184-194).
 * pred/targ [n_img, C, hw] with C = 10 (single layer, depth channels D = 8:10) or 20 (D = 16:20, :137);
 * rgba [n_img, rgba_c >= 4, hw] fp32, mask1 = rgba[:, 3] > 0 (:133). out8 (device): {total, depth_silog, colour term,
 * all_mse, all_mae, fb_divergence, #mask1 pixels, #maskD elements}. grad nullable, dtype of pred, scaled by grad_scale.
 * The SSIM and compositing terms (:168-183) belong to stdepth_utils.py and are out of scope.
 */
int mde_stdepth_loss(const void* pred, int pred_dtype, const float* targ, const float* rgba, int64_t rgba_c,
                     int64_t n_img, int64_t C, int64_t hw, int flags, float depth_w, float fbdiv_w,
                     float variance_focus, float grad_scale, void* ws, float* out8, void* grad, void* stream);

/* ---- depth -> point cloud ----------------------------------------------------------------- */
/*
 * point_cloud(depth, cam) (reference depth2pointcloud.py:12-31) for a batch of depth maps, with the
 * camera->world transform of :103-108 fused when matrix_world (16 host floats, row-major 4x4) is
 * non-null. out [n_img,h,w,3] in out_dtype64 ? fp64 (the reference's numpy promotion) : fp32.
 * Invalid pixels (depth outside (clip_start, clip_end)) give (0,0,NaN) before the transform.
 */
int mde_point_cloud(const float* depth, int64_t n_img, int64_t h, int64_t w, float angle_x,
                    float clip_start, float clip_end, const float* matrix_world_host, int out_f64,
                    void* out, void* stream);

/* ASCII PLY writer of reference depth2pointcloud.py:131-154 (host code; pointers are HOST memory): for every pixel
 * the front point then the back point (back_xyz nullable), each skipped when its x is NaN, as
 * "%f %f %f %d %d %d 0" with the cv2 BGR colour swapped to RGB; header, lines and the template's final newline.
 * front_xyz/back_xyz [n_points,3] float64 (world coordinates), color_bgr [n_points,3] uint8.
 * Returns the number of vertices written (>= 0) or a negative MDE_E* code. */
int64_t mde_write_ply(const char* path, const double* front_xyz, const double* back_xyz, const uint8_t* color_bgr,
                      int64_t n_points);

/* ---- misc ---------------------------------------------------------------------------------- */
size_t mde_workspace_bytes(int64_t max_images);
/* zero-fills the workspace and records its capacity (max_images); call once after allocating */
int mde_workspace_init(void* ws, int64_t max_images, void* stream);
const char* mde_last_error(void);
const char* mde_version(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
uint64_t mde_launch_count(void);
/* Profiling aid: device buffer of 8 x uint64 per CTA (>= 8 * 2 * SM count entries) that the masked-loss
 * kernels fill with GPU global-timer stamps at their phase boundaries (0 start, 1 reduce loop done,
 * 2 sums published, 3 grid barrier passed, 4 totals read, 5 gradient written); NULL disarms it. */
int mde_debug_set_trace(void* device_buf);
/* SM count / max co-resident CTAs the cooperative kernels use on the current device */
int mde_device_info(int* sm_count, int* coop_ctas);

#ifdef __cplusplus
}
#endif
#endif /* MDE_B200_H */
