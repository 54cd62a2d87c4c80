/*
 * include/redsec_b200.h -- C-ABI of the B200-native TFHE bootstrap engine behind REDsec's lib/ API.
 *
 * This is the drop-in boundary ("B-inner"): the entry points a REDsec maintainer binds
 * instead of lib/GPU/gates.cu + -lredcufhe.  Plain C, opaque handle, raw pointers and sizes; no torch
 * or C++ types.  Every call is stream-ordered on the context's stream; return value 0 = RS_OK,
 * otherwise rs_last_error() describes the failure.  There is no CPU fallback: without a CUDA device
 * rs_ctx_create fails.
 *
 * Ciphertext formats
 *   wire / host  : LWE sample = uint32[351]  (a[0..349], b)  -- TFHE LweSample order, torus32
 *   device       : rows of RS_LWE_STRIDE=352 words (word 351 = 0) so rows are 16-byte aligned
 * Keys (host, torus32):
 *   bsk[n][2l][2][N]      TGSW(s_i) rows: row r = c*l + p (c = input poly, p = level), 2 output polys
 *   ksk[N][t][base][n+1]  KS[i][j][h] = LWE(h * s'_i / base^(j+1))
 *
 * Each function cites the reference interface it replaces (paths relative to the REDsec tree).
 */
#ifndef REDSEC_B200_H
#define REDSEC_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RS_LWE_N 350
#define RS_LWE_WORDS 351
#define RS_LWE_STRIDE 352
#define RS_TLWE_N 1024
#define RS_BK_L 10
#define RS_BK_BGBIT 3
#define RS_KS_T 9
#define RS_KS_BASEBIT 3
#define RS_BSK_WORDS ((size_t)RS_LWE_N * 2 * RS_BK_L * 2 * RS_TLWE_N)
#define RS_KSK_WORDS ((size_t)RS_TLWE_N * RS_KS_T * 8 * RS_LWE_WORDS)
#define RS_EXT_STRIDE 1028

enum { RS_OK = 0, RS_ERR_CUDA = 1, RS_ERR_ARG = 2, RS_ERR_STATE = 3 };
/* gate ids: lib/GPU/gates.cu:246-286 (bootsNAND/OR/AND/NOR/XOR/XNOR) */
enum { RS_GATE_NAND = 0, RS_GATE_OR = 1, RS_GATE_AND = 2, RS_GATE_NOR = 3, RS_GATE_XOR = 4, RS_GATE_XNOR = 5 };
/* kernel kinds for rs_profile_get */
enum { RS_K_BLIND_ROTATE = 0, RS_K_KEYSWITCH = 1, RS_K_LINEAR = 2, RS_K_OTHER = 3, RS_K_COUNT = 4 };

typedef struct rs_ctx rs_ctx;

/* -- context ------------------------------------------------------------------------------------------
 * replaces redcufhe::SetGPUNum/Initialize(PubKey&) per device (nets/mnist/sign1024x1/net.cu:43-49),
 * redcufhe::CleanUp() (main.cu:84), Synchronize()/CuCheckError() (main.cu:74-75). */
int rs_ctx_create(rs_ctx **out, int device);
/* fails with RS_ERR_STATE while layer / net / communicator objects created on the context are still alive (they free their
 * device tables through it): destroy those first.  Every entry point makes the context's device current for the duration of
 * the call and restores the caller's, so one process may hold contexts on several devices.  Device-wide side effect of
 * rs_ctx_create: it raises cudaLimitPersistingL2CacheSize to the device maximum (the BSK stream uses an L2 evict_last hint). */
int rs_ctx_destroy(rs_ctx *ctx);
int rs_ctx_retain(rs_ctx *ctx);                      /* taken by redsec::Layer / rs_net / rs_comm for their lifetime */
int rs_ctx_release(rs_ctx *ctx);
int rs_pool_trim(rs_ctx *ctx);                       /* return the caching allocator's parked blocks to the driver (syncs) */
const char *rs_last_error(const rs_ctx *ctx);       /* ctx may be NULL: last creation error */
int rs_set_stream(rs_ctx *ctx, void *cuda_stream);  /* adopt a caller-owned cudaStream_t (NULL = own stream) */
int rs_get_stream(rs_ctx *ctx, void **cuda_stream); /* the cudaStream_t of the selected lane */
int rs_ctx_device(const rs_ctx *ctx);
int rs_sync(rs_ctx *ctx);                           /* waits for every lane */
/* Lanes: extra streams with their own bootstrap scratch, for independent chains of launches that should overlap (the
 * reference round-robins its per-ciphertext launches over 40 streams, lib/GPU/BinFunc_gpu.cu:116-138,599-621; here a lane
 * carries whole batched launches).  Lane 0 is the context's stream.  rs_lane_select(k) routes the following calls to lane k;
 * rs_lane_fork makes lanes 1.. wait for what lane 0 has issued so far, rs_lane_join makes lane 0 wait for lanes 1...
 * Buffers shared between lanes must be allocated before the fork and freed after the join. */
int rs_lanes(rs_ctx *ctx, int n);
int rs_lane_count(const rs_ctx *ctx);
int rs_lane_select(rs_ctx *ctx, int lane);
int rs_reserve_scratch(rs_ctx *ctx, size_t count);   /* size the selected lane's bootstrap scratch for `count` ciphertexts now (no allocation later) */
int rs_lane_fork(rs_ctx *ctx);
int rs_lane_join(rs_ctx *ctx);

/* -- evaluation key: north-star item (c) -------------------------------------------------------------
 * replaces new_tfheGateBootstrappingCloudKeySet_fromFile -> bk->bkFFT (nets/mnist/sign1024x1/net.cpp:53-55)
 * and ReadPubKeyFromFile + Initialize (net.cu:43-49).  Converts the BSK to the device-resident FP64
 * Fourier layout with a CUDA kernel and the KSK to the padded device table. */
int rs_load_eval_key(rs_ctx *ctx, const uint32_t *bsk_host, const uint32_t *ksk_host);

/* -- device LWE arrays: replaces tBit/tMultiBit arrays + CtxtCopyH2D/D2H (lib/GPU/gates.cu:9-23),
 * bit_calloc/mbit_calloc (lib/GPU/Layer.cu) */
int rs_lwe_alloc(rs_ctx *ctx, size_t count, uint32_t **dev_out);
int rs_lwe_free(rs_ctx *ctx, uint32_t *dev);
int rs_lwe_upload(rs_ctx *ctx, uint32_t *dev, const uint32_t *host_wire, size_t count);
int rs_lwe_download(rs_ctx *ctx, uint32_t *host_wire, const uint32_t *dev, size_t count);
int rs_lwe_copy(rs_ctx *ctx, uint32_t *dst_dev, const uint32_t *src_dev, size_t count);   /* device to device, stream ordered */
int rs_host_alloc(void **out, size_t bytes);        /* pinned host memory for the *_host entry points */
int rs_host_free(void *p);

/* -- the hot path ------------------------------------------------------------------------------------
 * rs_pbs_batch: count independent programmable bootstraps in one launch,
 *   out[c] = LWE(+mu if phase(in[c]) in [0,1/2) else -mu)
 * replaces BinOps::binarize_int / unbinarize_int -> tfhe_bootstrap_FFT (lib/BinOps_enc.cpp:182-192) and
 * redsec_binarize_bootstrap / redsec_unbinarize_bootstrap (lib/GPU/gates.cu:124-144), one call per layer
 * instead of one per neuron.  in == out is allowed. */
int rs_pbs_batch(rs_ctx *ctx, uint32_t *out_dev, const uint32_t *in_dev, size_t count, uint32_t mu);
/* rs_pbs_lut_batch: programmable bootstrap with caller-supplied test vectors, lut_dev = uint32[lut_mod][1024] on the
 * device; ciphertext c uses row c % lut_mod (rows are channel-fastest, so lut_mod = channels gives one table per channel):
 *   out[c] = LWE(+lut[j]) if phase(in[c]) rounds to j/2048 in [0,1/2), LWE(-lut[j-1024]) in [1/2,1).
 * The correct encrypted form of the DoReFa ReLU (multiply by slope, add bias, shift, clamp) that
 * {Int,Bin}Func::Quantize::relu_shift (lib/IntFunc.cpp:934-973, lib/BinFunc.cpp:1120-1162) attempts with
 * multiply_pc_ints + binarize_int + bootsMUX (functionally broken in the reference): ONE bootstrap per neuron
 * whose test vector is the per-channel staircase. */
int rs_pbs_lut_batch(rs_ctx *ctx, uint32_t *out_dev, const uint32_t *in_dev, size_t count, const uint32_t *lut_dev, int lut_mod);
/* rs_gate_batch: out[c] = boot((0,fix) +/- in0[c] +/- in1[c], mu); replaces bootsNAND..bootsXNOR
 * (lib/GPU/gates.cu:246-286) and BinOps::max -> bootsOR (lib/BinOps_enc.cpp:164-167). */
int rs_gate_batch(rs_ctx *ctx, int gate, uint32_t *out_dev, const uint32_t *in0_dev, const uint32_t *in1_dev,
                  size_t count, uint32_t mu);
/* host-buffer variants (wire format, H2D + compute + D2H inside the call): the reference-facing call a
 * per-ciphertext caller such as lib/GPU/BinFunc_gpu.cu:599-621 would make once per layer. */
int rs_pbs_batch_host(rs_ctx *ctx, uint32_t *out_host, const uint32_t *in_host, size_t count, uint32_t mu);
int rs_gate_batch_host(rs_ctx *ctx, int gate, uint32_t *out_host, const uint32_t *in0_host, const uint32_t *in1_host,
                       size_t count, uint32_t mu);

/* the two halves of rs_pbs_batch, exposed so sample-extract and keyswitch can be checked bit-exactly
 * on identical inputs: ext rows are uint32[RS_EXT_STRIDE] (a'[0..1023], b', pad) */
int rs_blind_rotate_batch(rs_ctx *ctx, uint32_t *ext_dev, const uint32_t *in_dev, size_t count, uint32_t mu);
int rs_keyswitch_batch(rs_ctx *ctx, uint32_t *out_dev, const uint32_t *ext_dev, size_t count);
int rs_ext_alloc(rs_ctx *ctx, size_t count, uint32_t **dev_out);
int rs_ext_upload(rs_ctx *ctx, uint32_t *dev, const uint32_t *host /*[count][1025]*/, size_t count);
int rs_ext_download(rs_ctx *ctx, uint32_t *host /*[count][1025]*/, const uint32_t *dev, size_t count);

/* -- bootstrap-free neighbours ---------------------------------------------------
 * out[o] = (0,bias[o]) + sum_k sign[k]*in[col[k]] over CSR row o; replaces the lweAddTo/lweSubTo loops of
 * {Bin,Int}Func::Convolution/SumPooling::execute and Quantize::execute's bias add
 * (lib/BinFunc.cpp:142-330,677-732,1044-1107; lib/IntFunc.cpp:152-319,643-700) and AddOp/SubOp
 * (lib/GPU/gates.cu:158-202).  rowptr/col/sign/bias are DEVICE pointers (uploaded once in prep). */
int rs_lwe_lincomb(rs_ctx *ctx, uint32_t *out_dev, size_t out_count, const uint32_t *in_dev, const int32_t *rowptr_dev,
                   const int32_t *col_dev, const int8_t *sign_dev, const uint32_t *bias_dev /* may be NULL */);
/* ternary convolution / fully-connected layer on LWE rows; replaces {Bin,Int}Func::Convolution::execute
 * (lib/BinFunc.cpp:142-330, lib/IntFunc.cpp:152-319; GPU twin lib/GPU/BinFunc_gpu.cu:110-222).  Activations are
 * (h,w,c) c-fastest (lib/BinFunc.cpp:404).  wpacked_dev: weights in {-1,0,+1} packed [ceil(out_dep/16)][K][16] int8 with
 * K index (fh*win_w+fw)*in_dep+di (the reference filter index lib/BinFunc.cpp:388-391 is k*OutDepth+od).
 * int_mode=1: zero weights and padding contribute -1/4096 (lib/IntFunc.cpp:268,277) instead of 0.
 * [od_begin,od_end) selects the output-channel slice this call computes (neuron sharding);
 * out rows are [pixel][od-od_begin]. */
typedef struct rs_conv_desc {
    int32_t in_h, in_w, in_dep;
    int32_t out_h, out_w, out_dep;
    int32_t win_h, win_w, stride_h, stride_w, ofs_h, ofs_w;
    int32_t int_mode;
    int32_t od_begin, od_end;
} rs_conv_desc;
int rs_lwe_conv(rs_ctx *ctx, uint32_t *out_dev, const uint32_t *in_dev, const int8_t *wpacked_dev,
                const uint32_t *bias_dev /* [out_dep] torus32, may be NULL */, const rs_conv_desc *desc);
/* dev[c].b += value for every row (adds the trivial sample (0,value)); lweNoiselessTrivial + lweAddTo, lib/BinOps_enc.cpp:137-141 */
int rs_lwe_add_const(rs_ctx *ctx, uint32_t *dev, size_t count, uint32_t value);
/* out = (0,fix) + m0*in0 + m1*in1 (in1 may be NULL; in place allowed): add_int / sub_int / mul_int / levelNOT of
 * lib/GPU/gates.cu:110-122,158-212 (AddOp / SubOp / NotOp) and lweAddTo / lweSubTo / lweAddMulTo for a batch */
int rs_lwe_axpby(rs_ctx *ctx, uint32_t *out_dev, const uint32_t *in0_dev, const uint32_t *in1_dev, size_t count, uint32_t m0,
                 uint32_t m1, uint32_t fix);
/* dev[r].b += bias_dev[r % mod] (per-channel bias, rows channel-fastest): the bias add of {Bin,Int}Func::Quantize::execute /
 * add_bias (lib/BinFunc.cpp:1063-1065,1085-1107) as a stage of its own for callers that compose Func objects */
int rs_lwe_add_bias(rs_ctx *ctx, uint32_t *dev, size_t count, const uint32_t *bias_dev, int mod);
int rs_dev_alloc(rs_ctx *ctx, size_t bytes, void **dev_out);
int rs_dev_free(rs_ctx *ctx, void *dev);
int rs_dev_upload(rs_ctx *ctx, void *dev, const void *host, size_t bytes);
int rs_dev_download(rs_ctx *ctx, void *host, const void *dev, size_t bytes);

/* restore canonical (pixel, channel) order after an all-gather of per-rank channel slices:
 * in [world][pixels][c_local]  ->  out [pixels][world*c_local]   (SURVEY 8e exchange step) */
int rs_lwe_interleave(rs_ctx *ctx, uint32_t *out_dev, const uint32_t *gathered_dev, size_t pixels, int c_local, int world);

/* -- exchange step between layers: NCCL all-gather of the ranks' output-channel blocks over NVLink, issued on the
 * context's stream (no host synchronisation).  The reference has none: with NUM_GPUS > 1 (lib/GPU/Layer.cuh:15) each OpenMP
 * thread fills only its own iterations of its own replica enc_segs[idx] (lib/GPU/BinFunc_gpu.cu:599-621) and nothing merges
 * them.  NCCL is bound at run time (dlopen libnccl.so.2); a single-GPU caller never needs it.
 *   rs_comm_init_rank: one process per GPU; rank 0 makes the id with rs_comm_unique_id and the caller distributes it
 *   rs_comm_init_all : one process, n contexts on n devices (the reference's one-host-thread-per-GPU model); each host thread
 *                      then drives its own context + communicator */
#define RS_COMM_ID_BYTES 128
typedef struct rs_comm rs_comm;
int rs_comm_unique_id(uint8_t *id /*[RS_COMM_ID_BYTES]*/);
int rs_comm_init_rank(rs_ctx *ctx, rs_comm **out, const uint8_t *id, int rank, int world);
int rs_comm_init_all(rs_ctx **ctxs, int n, rs_comm **comms_out /*[n]*/);
int rs_comm_destroy(rs_comm *comm);
int rs_comm_rank(const rs_comm *comm);
int rs_comm_world(const rs_comm *comm);
rs_ctx *rs_comm_ctx(const rs_comm *comm);
const char *rs_comm_last_error(void);
/* out[world][rows_per_rank] <- every rank's in[rows_per_rank] (device rows of RS_LWE_STRIDE words) */
int rs_allgather(rs_comm *comm, uint32_t *out_dev, const uint32_t *in_dev, size_t rows_per_rank);

/* -- client side (host C++) ------------------------------------------------------------
 * replaces client/gen_secure_keyset.cpp:94-120, client/encrypt_image.cpp:65-85, client/decrypt_image.cpp:46-63 */
uint32_t rs_modswitch_to_torus32(int32_t mu, int32_t msize);
int32_t rs_modswitch_from_torus32(uint32_t phase, int32_t msize);
/* Production entry points: randomness = ChaCha20 keyed with 256 bits from the OS (getrandom), a fresh key per call. */
int rs_keygen_secure(int32_t *lwe_key /*[350]*/, int32_t *tlwe_key /*[1024]*/, uint32_t *bsk, uint32_t *ksk);
int rs_lwe_encrypt_secure(uint32_t *ct_wire, const uint32_t *mu, size_t count, double alpha, const int32_t *lwe_key);
/* DETERMINISTIC variants for tests and known-answer vectors only (xoshiro256**, not a CSPRNG): the same seed gives the same
 * key / the same masks, so re-using a seed for two encryptions leaks the difference of their plaintexts. */
int rs_keygen(uint64_t seed, int32_t *lwe_key /*[350]*/, int32_t *tlwe_key /*[1024]*/, uint32_t *bsk, uint32_t *ksk);
int rs_lwe_encrypt(uint32_t *ct_wire, const uint32_t *mu, size_t count, double alpha, const int32_t *lwe_key, uint64_t seed);
int rs_selftest_chacha20(const uint32_t *key /*[8]*/, uint64_t domain, uint64_t index, uint32_t *out, size_t words);   /* KAT hook */
int rs_lwe_phase(uint32_t *phase, const uint32_t *ct_wire, size_t count, const int32_t *lwe_key);
int rs_lwe_decrypt(int32_t *msg, const uint32_t *ct_wire, size_t count, const int32_t *lwe_key, int32_t msize);
int rs_write_secret_key(const char *path, const int32_t *lwe_key, const int32_t *tlwe_key);
int rs_read_secret_key(const char *path, int32_t *lwe_key, int32_t *tlwe_key);
int rs_write_eval_key(const char *path, const uint32_t *bsk, const uint32_t *ksk);
int rs_read_eval_key(const char *path, uint32_t *bsk, uint32_t *ksk);
int rs_write_ctxt(const char *path, const uint32_t *ct_wire, size_t count, double variance, int append);
int rs_read_ctxt(const char *path, uint32_t *ct_wire, size_t count);

/* -- Layer forward: flat C view of the C++ classes in redsec_b200/host/redsec_layers.hpp ---
 * which mirror IntLayer/BinLayer (lib/GPU/IntLayer.cuh:16-34, lib/GPU/BinLayer.cuh:16-34) and the generated
 * HeBNN::init/run (nets/mnist/sign1024x1/net.cu).  Enum values are those of lib/Layer.h:58-101. */
typedef struct rs_net rs_net;
typedef struct rs_layer_params {    /* tNetParams, lib/Layer.h:126-167 */
    int32_t conv_win_h, conv_win_w, conv_stride_h, conv_stride_w, conv_same_pad;
    int32_t pool_win_h, pool_win_w, pool_stride_h, pool_stride_w, pool_same_pad;
    int32_t e_bias, shift_bits, version;
} rs_layer_params;
/* neuron partition of one layer: output-channel block [ch_begin,ch_end) of `rank`; whole layer when not shardable */
int rs_shard_range(int channels, int has_conv, int rank, int world, int *ch_begin, int *ch_end);
rs_net *rs_net_create(rs_ctx *ctx);     /* ctx may be NULL: a net that can be prepared and asked for shapes / shard plan, not run */
void rs_net_destroy(rs_net *net);
int rs_net_add_layer(rs_net *net, int int_layer, int conv_type, int out_depth, int pool_type, int quant_type,
                     const rs_layer_params *p);
int rs_net_prep(rs_net *net, const char *weights_path /* var_prep.dat */, int in_h, int in_w, int in_dep);
/* same with the input tDimensions of the generated net spelled out (lay_dim.in_bits / up_bound / scale,
 * nets/mnist/relu1024x1/net.cpp:96-110: 2, 2, 1); rs_net_prep uses the sign nets' 9, 510, 255 */
int rs_net_prep_ex(rs_net *net, const char *weights_path, int in_h, int in_w, int in_dep, int in_bits, int up_bound, float scale);
/* optional, after rs_net_prep: build and upload the device tables (packed weights, pooling rows, OR trees, ReLU test vectors) of
 * rank's channel slices now rather than lazily inside the first forward -- the counterpart of the weight loading that
 * {Bin,Int}Layer::prep does before the reference's "Inference Time" window (nets/mnist/sign1024x1/main.cu:72-78) */
int rs_net_build_tables(rs_net *net, int rank, int world);
int rs_net_num_layers(const rs_net *net);
int rs_net_layer_info(rs_net *net, int layer, size_t *out_count, int *channels, size_t *bootstraps, int *out_h, int *out_w);
/* forward of one layer for rank's output-channel slice (world=1: whole layer).  Does not free in_dev; the caller
 * frees *out_dev with rs_lwe_free.  Output rows are [pixel][ch_begin..ch_end). */
int rs_net_layer_forward(rs_net *net, int layer, const uint32_t *in_dev, size_t in_count, int rank, int world,
                         uint32_t **out_dev, size_t *out_count, int *ch_begin, int *ch_end);

/* how layer `layer` is split over `world` ranks: *mode 0 = computed whole on every rank, 1 = output-channel blocks
 * (all-gather + rs_lwe_interleave), 2 = output-pixel blocks (all-gather only); *rows_per_rank = ciphertexts each rank
 * contributes to the all-gather; *c_local = channels per rank block.  Host logic only (works on a net created without ctx). */
int rs_net_shard_plan(rs_net *net, int layer, int world, int *mode, size_t *rows_per_rank, int *c_local);
/* one layer, neuron-sharded over the communicator (comm == NULL: whole layer on this GPU); does not consume in_dev */
int rs_net_layer_forward_sharded(rs_net *net, int layer, rs_comm *comm, const uint32_t *in_dev, size_t in_count,
                                 uint32_t **out_dev, size_t *out_count);
/* HeBNN::run (nets/mnist/sign1024x1/net.cu:105-120) for the whole network on the engine stream.  comm == NULL: one GPU.  Otherwise every layer is
 * neuron-sharded over the communicator's ranks (each GPU a full key replica, output channels partitioned) with an NCCL
 * all-gather between layers and NO host synchronisation; every rank receives the full output.  Does not consume in_dev;
 * the caller frees *out_dev with rs_lwe_free. */
int rs_net_run(rs_net *net, rs_comm *comm, const uint32_t *in_dev, size_t in_count, uint32_t **out_dev, size_t *out_count);

/* -- measurement ------------------------------------------------------------------------------------- */
int rs_profile_enable(rs_ctx *ctx, int on);         /* bracket every kernel launch with CUDA events on the ctx stream */
int rs_profile_get(rs_ctx *ctx, int kind, double *total_ms, uint64_t *launches);   /* syncs; since last reset */
int rs_profile_reset(rs_ctx *ctx);
uint64_t rs_launch_count(const rs_ctx *ctx);        /* kernels launched by this context since creation */
int rs_fp64_peak(rs_ctx *ctx, double *tflops);      /* dependent-free DFMA loop on all SMs: the FP64 roofline denominator */
int rs_fp64_peak_three_operand(rs_ctx *ctx, double *tflops);   /* same loop, three distinct register operands per FMA (register-file bound) */
int rs_set_tuning(rs_ctx *ctx, int br_variant);    /* blind-rotate kernel: 0 = warp-specialised (default); 1 / 2 = single-role, 7- / 4-stage BSK ring; 3 / 4 = warp-specialised with the BSK served from tensor memory (experimental) */
int rs_set_ks_variant(rs_ctx *ctx, int ks_variant);   /* keyswitch kernel: 0 = auto (default: exact int8 GEMM on the tensor cores, tcgen05.mma kind::i8, for batches >= 2048; shared-memory gather with a TMA-streamed key below), 1 = un-tiled L1/L2 gather, 2 = tensor cores always, 3 = shared-memory gather always */
int rs_device_info(rs_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor, size_t *smem_optin);

#ifdef __cplusplus
}
#endif
#endif
