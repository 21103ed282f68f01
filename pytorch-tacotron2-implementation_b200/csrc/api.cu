// C ABI of libb200taco (include/b200taco.h): argument checking, workspace carving and the
// launch sequences of the encoder / decoder / postnet stages.  Nothing throws; every entry
// point returns a b200taco_status and records a message for b200taco_last_error().
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>

#include "handle.h"

namespace b200taco {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int64_t& launch_counter() { return g_launches; }

namespace {

// teacher forcing: decoder input frames, shifted right by one with a zero go-frame
// (src/model.py:281-285).  X[s*B + b][o] = s == 0 ? 0 : targets[b][o][s-1]
__global__ void tf_frames_kernel(const float* __restrict__ targets, float* __restrict__ X, int B, int S) {
    const size_t total = (size_t)S * B * N_MELS;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int o = i % N_MELS, b = (i / N_MELS) % B, s = i / ((size_t)N_MELS * B);
        X[i] = s == 0 ? 0.f : targets[((size_t)b * N_MELS + o) * S + (s - 1)];
    }
}
struct Carver {
    uint8_t* base;
    size_t off = 0;
    float* f(size_t n) {
        float* p = base ? reinterpret_cast<float*>(base + off) : nullptr;
        off = align_up(off + n * sizeof(float), 256);
        return p;
    }
    void* raw(size_t bytes) {
        void* p = base ? base + off : nullptr;
        off = align_up(off + bytes, 256);
        return p;
    }
};

struct EncoderWs { float *x0, *x1, *xproj; size_t total; };
EncoderWs encoder_ws(void* base, int B, int T) {
    Carver c{static_cast<uint8_t*>(base)};
    EncoderWs w{};
    w.x0 = c.f((size_t)B * T * EMB);
    w.x1 = c.f((size_t)B * T * EMB);
    w.xproj = c.f((size_t)B * T * 8 * ENC_H);
    w.total = c.off;
    return w;
}
struct DecoderWs { float *pm, *frames, *y1, *y2, *p_all; void* scratch; size_t scratch_bytes, total; };
DecoderWs decoder_ws(void* base, int B, int T, int S_tf) {
    Carver c{static_cast<uint8_t*>(base)};
    DecoderWs w{};
    w.pm = c.f((size_t)B * T * ATT_DIM);
    w.scratch_bytes = decoder_scratch_bytes(B, T);
    w.scratch = c.raw(w.scratch_bytes);
    if (S_tf > 0) {
        w.frames = c.f((size_t)S_tf * B * N_MELS);
        w.y1 = c.f((size_t)S_tf * B * PRENET);
        w.y2 = c.f((size_t)S_tf * B * PRENET);
        w.p_all = c.f((size_t)S_tf * PRENET * (decoder_batch_pad(B) < 32 ? 32 : decoder_batch_pad(B)));
    }
    w.total = c.off;
    return w;
}
struct PostnetWs { float *a, *b; size_t total; };
PostnetWs postnet_ws(void* base, int B, int S) {
    Carver c{static_cast<uint8_t*>(base)};
    PostnetWs w{};
    w.a = c.f((size_t)B * S * POST_DIM);
    w.b = c.f((size_t)B * S * POST_DIM);
    w.total = c.off;
    return w;
}

int check_handle(const b200taco_handle* h, bool need_packed) {
    TACO_REQUIRE(h != nullptr, B200TACO_ERR_INVALID, "null handle");
    if (need_packed)
        TACO_REQUIRE(h->packed, B200TACO_ERR_STATE, "weights not packed: call b200taco_pack_weights first");
    TACO_CUDA(cudaSetDevice(h->device));
    return B200TACO_OK;
}

int check_shape(int B, int T) {
    TACO_REQUIRE(B >= 1 && B <= MAX_BATCH, B200TACO_ERR_UNSUPPORTED, "batch %d outside [1, %d]", B, MAX_BATCH);
    TACO_REQUIRE(T >= 1 && T <= MAX_TEXT_LEN, B200TACO_ERR_UNSUPPORTED, "text length %d outside [1, %d]", T,
                 MAX_TEXT_LEN);
    return B200TACO_OK;
}

}  // namespace
}  // namespace b200taco

using namespace b200taco;

#define TACO_TRY(expr)                   \
    do {                                 \
        int _s = (expr);                 \
        if (_s != B200TACO_OK) return _s; \
    } while (0)

extern "C" {

B200TACO_API int32_t b200taco_abi_version(void) { return B200TACO_ABI_VERSION; }
B200TACO_API const char* b200taco_last_error(void) { return g_err; }
B200TACO_API int32_t b200taco_max_batch(void) { return MAX_BATCH; }
B200TACO_API int32_t b200taco_max_text_len(void) { return MAX_TEXT_LEN; }

B200TACO_API int32_t b200taco_create(const b200taco_hparams* hp, int32_t device, b200taco_handle** out) {
    TACO_REQUIRE(hp && out, B200TACO_ERR_INVALID, "b200taco_create: null argument");
    *out = nullptr;
    const b200taco_hparams want = {N_SYMBOLS, EMB, ENC_CONVS, ENC_KERNEL, N_MELS, RNN, PRENET, RNN, ATT_DIM,
                                   LOC_F, LOC_K, POST_DIM, POST_KERNEL, POST_CONVS, hp->max_decoder_steps};
    TACO_REQUIRE(memcmp(&want, hp, sizeof(want)) == 0 && hp->max_decoder_steps >= 1, B200TACO_ERR_HPARAMS,
                 "hparams differ from the compiled specialisation (symbols %d, emb %d, convs %dx%d, mels %d, rnn %d, "
                 "prenet %d, attention %d/%d/%d, postnet %d/%d/%d)",
                 N_SYMBOLS, EMB, ENC_CONVS, ENC_KERNEL, N_MELS, RNN, PRENET, ATT_DIM, LOC_F, LOC_K, POST_DIM,
                 POST_KERNEL, POST_CONVS);
    int count = 0;
    TACO_CUDA(cudaGetDeviceCount(&count));
    TACO_REQUIRE(device >= 0 && device < count, B200TACO_ERR_CUDA, "no CUDA device %d (found %d); there is no CPU fallback",
                 device, count);
    TACO_CUDA(cudaSetDevice(device));
    b200taco_handle* h = new (std::nothrow) b200taco_handle();
    TACO_REQUIRE(h, B200TACO_ERR_STATE, "out of host memory");
    h->device = device;
    cudaError_t e = cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device);
    int coop = 0, major = 0;
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    if (e != cudaSuccess || !coop || major != 10) {
        delete h;
        set_error("device %d unusable: %s, cooperative launch %d, compute capability major %d (need sm_100a)", device,
                  cudaGetErrorString(e), coop, major);
        return B200TACO_ERR_CUDA;
    }
    *out = h;
    return B200TACO_OK;
}

B200TACO_API int32_t b200taco_destroy(b200taco_handle* h) {
    if (!h) return B200TACO_OK;
    cudaSetDevice(h->device);
    if (h->arena) cudaFree(h->arena);
    delete h;
    return B200TACO_OK;
}

B200TACO_API int32_t b200taco_set_weight_dtype(b200taco_handle* h, int32_t dtype) {
    TACO_REQUIRE(h != nullptr, B200TACO_ERR_INVALID, "null handle");
    TACO_REQUIRE(dtype == B200TACO_WEIGHTS_BF16 || dtype == B200TACO_WEIGHTS_FP32, B200TACO_ERR_INVALID,
                 "unknown weight dtype %d", dtype);
    if (h->weight_dtype != dtype) h->packed = false;
    h->weight_dtype = dtype;
    return B200TACO_OK;
}

B200TACO_API int32_t b200taco_set_decoder_path(b200taco_handle* h, int32_t path) {
    TACO_REQUIRE(h != nullptr, B200TACO_ERR_INVALID, "null handle");
    TACO_REQUIRE(path == B200TACO_PATH_AUTO || path == B200TACO_PATH_SIMT, B200TACO_ERR_INVALID, "unknown decoder path %d", path);
    h->decoder_path = path;
    return B200TACO_OK;
}

B200TACO_API int32_t b200taco_set_dense_path(b200taco_handle* h, int32_t path) {
    TACO_REQUIRE(h != nullptr, B200TACO_ERR_INVALID, "null handle");
    TACO_REQUIRE(path == B200TACO_PATH_AUTO || path == B200TACO_PATH_SIMT || path == B200TACO_PATH_TC, B200TACO_ERR_INVALID,
                 "unknown dense path %d", path);
    h->dense_path = path;
    return B200TACO_OK;
}

B200TACO_API int32_t b200taco_pack_weights(b200taco_handle* h, const b200taco_weights* w, void* stream) {
    TACO_TRY(check_handle(h, false));
    TACO_REQUIRE(w != nullptr, B200TACO_ERR_INVALID, "null weights");
    const void* const* ptrs = reinterpret_cast<const void* const*>(w);
    for (size_t i = 0; i < sizeof(b200taco_weights) / sizeof(void*); ++i)
        TACO_REQUIRE(ptrs[i] != nullptr, B200TACO_ERR_INVALID, "weights field %zu is null", i);
    return pack_weights(h, w, static_cast<cudaStream_t>(stream));
}

B200TACO_API size_t b200taco_workspace_bytes(const b200taco_handle* h, int32_t B, int32_t T, int32_t S) {
    (void)h;
    if (B < 1 || T < 1 || S < 1) return 0;
    size_t need = encoder_ws(nullptr, B, T).total;
    size_t d = decoder_ws(nullptr, B, T, S).total + 1024;
    size_t p = postnet_ws(nullptr, B, S).total;
    need = d > need ? d : need;
    need = p > need ? p : need;
    return need;
}

static int encoder_run(b200taco_handle* h, const int64_t* d_ids, const int32_t* d_lengths, int B, int T, float* d_memory,
                       void* d_workspace, size_t workspace_bytes, cudaStream_t st) {
    const EncoderWs ws = encoder_ws(d_workspace, B, T);
    TACO_REQUIRE(workspace_bytes >= ws.total, B200TACO_ERR_STATE, "encoder workspace too small: %zu < %zu", workspace_bytes,
                 ws.total);
    const int n = B * T;
    TACO_TRY(embedding_gather(d_ids, h->enc_emb, ws.x0, n, st));
    // ragged batch: what lies behind an utterance's own length is ZERO at the input of every convolution, like the zero padding
    // of its exact-length solo run (the padded ids would otherwise leak into its last two positions per layer)
    ChannelMajorOut tail{};
    tail.lengths = d_lengths;
    tail.zero_tail = d_lengths != nullptr;
    if (d_lengths != nullptr) TACO_TRY(zero_tail_rows(ws.x0, EMB, d_lengths, B, T, EMB, st));
    float* src = ws.x0;
    float* dst = ws.x1;
    for (int i = 0; i < ENC_CONVS; ++i) {
        TACO_TRY(dense_layer(h, src, EMB, h->enc_conv_w[i], h->enc_conv_wtc[i], h->enc_conv_scale[i], h->enc_conv_shift[i], nullptr, 0,
                                      dst, EMB, n, T, EMB, EMB, ENC_KERNEL, ACT_RELU, true, st, tail));
        float* t = src; src = dst; dst = t;
    }
    TACO_TRY(dense_layer(h, src, EMB, h->enc_lstm_wih, h->enc_lstm_wih_tc, nullptr, h->enc_lstm_bias, nullptr, 0, ws.xproj, 8 * ENC_H, n,
                                  n, EMB, 8 * ENC_H, 1, ACT_NONE, true, st));
    TACO_TRY(bilstm_recurrence(h, ws.xproj, d_memory, B, T, st, d_lengths));
    if (d_lengths != nullptr) TACO_TRY(zero_tail_rows(d_memory, EMB, d_lengths, B, T, EMB, st));
    return B200TACO_OK;
}

B200TACO_API int32_t b200taco_encoder_fwd(b200taco_handle* h, const int64_t* d_ids, int32_t B, int32_t T, float* d_memory,
                             void* d_workspace, size_t workspace_bytes, void* stream) {
    TACO_TRY(check_handle(h, true));
    TACO_TRY(check_shape(B, T));
    TACO_REQUIRE(d_ids && d_memory && d_workspace, B200TACO_ERR_INVALID, "encoder_fwd: null pointer");
    return encoder_run(h, d_ids, nullptr, B, T, d_memory, d_workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

B200TACO_API int32_t b200taco_encoder_fwd_ragged(b200taco_handle* h, const int64_t* d_ids, const int32_t* d_lengths, int32_t B,
                                                 int32_t T, float* d_memory, void* d_workspace, size_t workspace_bytes,
                                                 void* stream) {
    TACO_TRY(check_handle(h, true));
    TACO_TRY(check_shape(B, T));
    TACO_REQUIRE(d_ids && d_lengths && d_memory && d_workspace, B200TACO_ERR_INVALID, "encoder_fwd_ragged: null pointer");
    return encoder_run(h, d_ids, d_lengths, B, T, d_memory, d_workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

static int run_decoder(b200taco_handle* h, int mode, const float* d_memory, const float* d_mel_targets,
                       const int64_t* d_text_lengths, int B, int T, int S, float thr, int stop_mode, float* d_mel,
                       float* d_gate, float* d_align, int32_t* d_lengths, int32_t* d_steps, void* d_workspace,
                       size_t workspace_bytes, cudaStream_t st) {
    const DecoderWs ws = decoder_ws(d_workspace, B, T, mode == 1 ? S : 0);
    TACO_REQUIRE(workspace_bytes >= ws.total, B200TACO_ERR_STATE, "decoder workspace too small: %zu < %zu", workspace_bytes,
                 ws.total);
    // processed_memory = memory_layer(memory), once per utterance (src/model.py:112)
    TACO_TRY(dense_layer(h, d_memory, EMB, h->attn_memory_wT, h->attn_memory_wtc, nullptr, nullptr, nullptr, 0, ws.pm, ATT_DIM, B * T,
                                  B * T, EMB, ATT_DIM, 1, ACT_NONE, true, st));
    if (mode == 1) {
        // all prenet outputs at once: the inputs of every step are known (teacher forcing)
        const int n = S * B;
        tf_frames_kernel<<<(int)(((size_t)n * N_MELS + 255) / 256 > 2048 ? 2048 : ((size_t)n * N_MELS + 255) / 256), 256, 0,
                           st>>>(d_mel_targets, ws.frames, B, S);
        TACO_LAUNCHED();
        TACO_TRY(dense_layer(h, ws.frames, N_MELS, h->prenet1T, h->prenet1_tc, nullptr, nullptr, nullptr, 0, ws.y1, PRENET, n, n,
                                      N_MELS, PRENET, 1, ACT_RELU, false, st));
        TACO_TRY(dense_layer(h, ws.y1, PRENET, h->prenet2T, h->prenet2_tc, nullptr, nullptr, nullptr, 0, ws.y2, PRENET, n, n, PRENET,
                                      PRENET, 1, ACT_RELU, false, st));
        TACO_TRY(decoder_pack_p_all(h, ws.y2, ws.p_all, B, S, st));
    }
    DecoderArgs a{};
    a.B = B; a.T = T; a.max_steps = S; a.mode = mode; a.stop_mode = stop_mode; a.gate_threshold = thr;
    a.memory = d_memory; a.pm = ws.pm; a.text_lengths = d_text_lengths; a.p_all = ws.p_all;
    a.mel = d_mel; a.gate = d_gate; a.align = d_align; a.lengths = d_lengths; a.steps_run = d_steps;
    a.scratch = ws.scratch; a.scratch_bytes = ws.scratch_bytes;
    return decoder_run(h, a, st);
}

B200TACO_API int32_t b200taco_decoder_infer(b200taco_handle* h, const float* d_memory, int32_t B, int32_t T, int32_t max_steps,
                               float gate_threshold, int32_t stop_mode, float* d_mel, float* d_gate, float* d_align,
                               int32_t* d_lengths, int32_t* d_steps_run, void* d_workspace, size_t workspace_bytes,
                               void* stream) {
    TACO_TRY(check_handle(h, true));
    TACO_TRY(check_shape(B, T));
    TACO_REQUIRE(d_memory && d_mel && d_gate && d_align && d_lengths && d_steps_run && d_workspace, B200TACO_ERR_INVALID,
                 "decoder_infer: null pointer");
    TACO_REQUIRE(max_steps >= 1, B200TACO_ERR_INVALID, "max_steps must be >= 1 (got %d)", max_steps);
    TACO_REQUIRE(stop_mode == B200TACO_STOP_BATCH_WIDE || stop_mode == B200TACO_STOP_PER_UTTERANCE, B200TACO_ERR_INVALID,
                 "unknown stop mode %d", stop_mode);
    return run_decoder(h, 0, d_memory, nullptr, nullptr, B, T, max_steps, gate_threshold, stop_mode, d_mel, d_gate, d_align,
                       d_lengths, d_steps_run, d_workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

// Utterance queue: processed memory of all N utterances, then the scratch of a `rows`-row launch.
static size_t stream_ws_bytes(int N, int rows, int T) {
    return align_up((size_t)N * T * ATT_DIM * sizeof(float), 256) + decoder_scratch_bytes(rows, T) + 256;
}

B200TACO_API size_t b200taco_stream_workspace_bytes(const b200taco_handle*, int32_t N, int32_t rows, int32_t T) {
    if (N < 1 || rows < 1 || rows > MAX_BATCH || T < 1 || T > MAX_TEXT_LEN) return 0;
    return stream_ws_bytes(N, rows, T);
}

static int32_t infer_stream(b200taco_handle* h, const float* d_memory, const int64_t* d_text_lengths, int32_t N, int32_t rows, int32_t T,
                            int32_t max_steps, float gate_threshold, float* d_mel, float* d_gate, float* d_align, int32_t* d_lengths,
                            int32_t* d_steps_run, void* d_workspace, size_t workspace_bytes, void* stream) {
    TACO_TRY(check_handle(h, true));
    TACO_TRY(check_shape(rows, T));
    TACO_REQUIRE(d_memory && d_mel && d_gate && d_align && d_lengths && d_steps_run && d_workspace, B200TACO_ERR_INVALID,
                 "decoder_infer_stream: null pointer");
    TACO_REQUIRE(max_steps >= 1, B200TACO_ERR_INVALID, "max_steps must be >= 1 (got %d)", max_steps);
    TACO_REQUIRE(N >= rows && N % rows == 0, B200TACO_ERR_INVALID,
                 "decoder_infer_stream: %d utterances is not a positive multiple of %d rows", N, rows);
    TACO_REQUIRE(workspace_bytes >= stream_ws_bytes(N, rows, T), B200TACO_ERR_STATE, "decoder workspace too small: %zu < %zu",
                 workspace_bytes, stream_ws_bytes(N, rows, T));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* pm = static_cast<float*>(d_workspace);
    uint8_t* scratch = static_cast<uint8_t*>(d_workspace) + align_up((size_t)N * T * ATT_DIM * sizeof(float), 256);
    // processed_memory = memory_layer(memory), once per utterance (src/model.py:112)
    TACO_TRY(dense_layer(h, d_memory, EMB, h->attn_memory_wT, h->attn_memory_wtc, nullptr, nullptr, nullptr, 0, pm, ATT_DIM, N * T,
                         N * T, EMB, ATT_DIM, 1, ACT_NONE, true, st));
    DecoderArgs a{};
    a.B = rows; a.T = T; a.max_steps = max_steps; a.mode = 0; a.stop_mode = B200TACO_STOP_PER_UTTERANCE;
    a.gate_threshold = gate_threshold; a.n_utts = N;
    a.memory = d_memory; a.pm = pm; a.text_lengths = d_text_lengths; a.p_all = nullptr;
    a.mel = d_mel; a.gate = d_gate; a.align = d_align; a.lengths = d_lengths; a.steps_run = d_steps_run;
    a.scratch = scratch; a.scratch_bytes = decoder_scratch_bytes(rows, T);
    return decoder_run(h, a, st);
}

B200TACO_API int32_t b200taco_decoder_infer_stream(b200taco_handle* h, const float* d_memory, int32_t N, int32_t rows, int32_t T,
                                                   int32_t max_steps, float gate_threshold, float* d_mel, float* d_gate,
                                                   float* d_align, int32_t* d_lengths, int32_t* d_steps_run, void* d_workspace,
                                                   size_t workspace_bytes, void* stream) {
    return infer_stream(h, d_memory, nullptr, N, rows, T, max_steps, gate_threshold, d_mel, d_gate, d_align, d_lengths, d_steps_run,
                        d_workspace, workspace_bytes, stream);
}

B200TACO_API int32_t b200taco_decoder_infer_stream_ragged(b200taco_handle* h, const float* d_memory, const int64_t* d_text_lengths,
                                                          int32_t N, int32_t rows, int32_t T, int32_t max_steps, float gate_threshold,
                                                          float* d_mel, float* d_gate, float* d_align, int32_t* d_lengths,
                                                          int32_t* d_steps_run, void* d_workspace, size_t workspace_bytes, void* stream) {
    TACO_REQUIRE(d_text_lengths != nullptr, B200TACO_ERR_INVALID, "decoder_infer_stream_ragged: null text lengths");
    return infer_stream(h, d_memory, d_text_lengths, N, rows, T, max_steps, gate_threshold, d_mel, d_gate, d_align, d_lengths, d_steps_run,
                        d_workspace, workspace_bytes, stream);
}

B200TACO_API int32_t b200taco_decoder_teacher_forced(b200taco_handle* h, const float* d_memory, const float* d_mel_targets,
                                        const int64_t* d_text_lengths, int32_t B, int32_t T, int32_t S, float* d_mel,
                                        float* d_gate, float* d_align, void* d_workspace, size_t workspace_bytes,
                                        void* stream) {
    TACO_TRY(check_handle(h, true));
    TACO_TRY(check_shape(B, T));
    TACO_REQUIRE(d_memory && d_mel_targets && d_mel && d_gate && d_align && d_workspace, B200TACO_ERR_INVALID,
                 "decoder_teacher_forced: null pointer");
    TACO_REQUIRE(S >= 1, B200TACO_ERR_INVALID, "S must be >= 1 (got %d)", S);
    // lengths/steps outputs of the kernel land in the control area of the workspace tail
    const DecoderWs ws = decoder_ws(d_workspace, B, T, S);
    TACO_REQUIRE(workspace_bytes >= ws.total + 1024, B200TACO_ERR_STATE, "decoder workspace too small: %zu < %zu",
                 workspace_bytes, ws.total + 1024);
    int32_t* tail = reinterpret_cast<int32_t*>(static_cast<uint8_t*>(d_workspace) + ws.total);
    return run_decoder(h, 1, d_memory, d_mel_targets, d_text_lengths, B, T, S, 2.0f, B200TACO_STOP_BATCH_WIDE, d_mel, d_gate,
                       d_align, tail, tail + MAX_BATCH, d_workspace, ws.total, static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------------------------------
// Teacher-forced decoder in TRAIN mode (Decoder.forward with self.training, src/model.py:277-297): the same persistent
// kernel, with dropout masks on the prenet outputs (p=0.5, :170) and on the carried LSTM hidden states (p=0.1, :262,
// :270), keeping per-step state for the backward pass (train_api.cu: b200taco_decoder_train_bwd).
// saved (floats): frames [S*B][80] | y1 [S*B][256] | y2 [S*B][256] | pm [B*T][128] | ga [S*B][4096] | ca [S*B][1024] |
//                 ah [S*B][1024] | gd [S*B][4096] | cd [S*B][1024] | dh [S*B][1024] | ctx [S*B][512]
}  // extern "C"
namespace b200taco {
DecSaved dec_saved(int B, int T, int S) {
    DecSaved d{};
    const size_t n = (size_t)S * B;
    size_t off = 0;
    d.frames = off; off += n * N_MELS;
    d.y1 = off; off += n * PRENET;
    d.y2 = off; off += n * PRENET;
    d.pm = off; off += (size_t)B * T * ATT_DIM;
    d.ga = off; off += n * 4 * RNN;
    d.ca = off; off += n * RNN;
    d.ah = off; off += n * RNN;
    d.gd = off; off += n * 4 * RNN;
    d.cd = off; off += n * RNN;
    d.dh = off; off += n * RNN;
    d.ctx = off; off += n * EMB;
    d.total = off;
    return d;
}
int relu_drop_fwd(float* y, const float* mask, float scale, size_t n, cudaStream_t st);
}  // namespace b200taco
extern "C" {

B200TACO_API size_t b200taco_decoder_train_saved_bytes(int32_t B, int32_t T, int32_t S) {
    return dec_saved(B, T, S).total * sizeof(float);
}

B200TACO_API int32_t b200taco_decoder_train_fwd(b200taco_handle* h, const float* d_memory, const float* d_mel_targets,
                                                const int64_t* d_text_lengths, int32_t B, int32_t T, int32_t S,
                                                const float* d_mask_p1, const float* d_mask_p2, float prenet_scale,
                                                const float* d_mask_a, const float* d_mask_d, float lstm_scale, float* d_mel,
                                                float* d_gate, float* d_align, float* d_saved, size_t saved_bytes,
                                                void* d_workspace, size_t workspace_bytes, void* stream) {
    TACO_TRY(check_handle(h, true));
    TACO_TRY(check_shape(B, T));
    TACO_REQUIRE(d_memory && d_mel_targets && d_mel && d_gate && d_align && d_saved && d_workspace && S >= 1, B200TACO_ERR_INVALID,
                 "decoder_train_fwd: bad argument");
    const DecSaved sv = dec_saved(B, T, S);
    TACO_REQUIRE(saved_bytes >= sv.total * sizeof(float), B200TACO_ERR_STATE, "decoder_train_fwd: saved buffer too small");
    const DecoderWs ws = decoder_ws(d_workspace, B, T, S);
    TACO_REQUIRE(workspace_bytes >= ws.total + 1024, B200TACO_ERR_STATE, "decoder workspace too small: %zu < %zu", workspace_bytes,
                 ws.total + 1024);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* frames = d_saved + sv.frames;
    float* y1 = d_saved + sv.y1;
    float* y2 = d_saved + sv.y2;
    float* pm = d_saved + sv.pm;
    const int n = S * B;
    // fp32 CUDA-core dense layers throughout the training path (exact round-to-nearest accumulation)
    TACO_TRY(conv1d_channels_last(d_memory, EMB, h->attn_memory_wT, nullptr, nullptr, nullptr, 0, pm, ATT_DIM, B * T, B * T, EMB,
                                  ATT_DIM, 1, ACT_NONE, st));
    tf_frames_kernel<<<(int)(((size_t)n * N_MELS + 255) / 256 > 2048 ? 2048 : ((size_t)n * N_MELS + 255) / 256), 256, 0, st>>>(
        d_mel_targets, frames, B, S);
    TACO_LAUNCHED();
    TACO_TRY(conv1d_channels_last(frames, N_MELS, h->prenet1T, nullptr, nullptr, nullptr, 0, y1, PRENET, n, n, N_MELS, PRENET, 1,
                                  ACT_RELU, st));
    TACO_TRY(relu_drop_fwd(y1, d_mask_p1, prenet_scale, (size_t)n * PRENET, st));
    TACO_TRY(conv1d_channels_last(y1, PRENET, h->prenet2T, nullptr, nullptr, nullptr, 0, y2, PRENET, n, n, PRENET, PRENET, 1,
                                  ACT_RELU, st));
    TACO_TRY(relu_drop_fwd(y2, d_mask_p2, prenet_scale, (size_t)n * PRENET, st));
    TACO_TRY(decoder_pack_p_all(h, y2, ws.p_all, B, S, st));
    int32_t* tail = reinterpret_cast<int32_t*>(static_cast<uint8_t*>(d_workspace) + ws.total);
    DecoderArgs a{};
    a.B = B; a.T = T; a.max_steps = S; a.mode = 1; a.stop_mode = B200TACO_STOP_BATCH_WIDE; a.gate_threshold = 2.0f;
    a.memory = d_memory; a.pm = pm; a.text_lengths = d_text_lengths; a.p_all = ws.p_all;
    a.mel = d_mel; a.gate = d_gate; a.align = d_align; a.lengths = tail; a.steps_run = tail + MAX_BATCH;
    a.scratch = ws.scratch; a.scratch_bytes = ws.scratch_bytes;
    a.sv_ga = d_saved + sv.ga; a.sv_ca = d_saved + sv.ca; a.sv_ah = d_saved + sv.ah;
    a.sv_gd = d_saved + sv.gd; a.sv_cd = d_saved + sv.cd; a.sv_dh = d_saved + sv.dh; a.sv_ctx = d_saved + sv.ctx;
    a.mask_a = d_mask_a; a.mask_d = d_mask_d; a.drop_scale = lstm_scale;
    return decoder_run(h, a, st);
}

static int postnet_run(b200taco_handle* h, const float* d_mel_coarse, int32_t B, int32_t S, float* d_mel_post, const ChannelMajorOut& cm,
                       void* d_workspace, size_t workspace_bytes, void* stream, const char* who) {
    TACO_TRY(check_handle(h, true));
    TACO_REQUIRE(d_mel_coarse && d_mel_post && d_workspace, B200TACO_ERR_INVALID, "%s: null pointer", who);
    TACO_REQUIRE(B >= 1 && S >= 1, B200TACO_ERR_INVALID, "%s: B %d, S %d", who, B, S);
    const PostnetWs ws = postnet_ws(d_workspace, B, S);
    TACO_REQUIRE(workspace_bytes >= ws.total, B200TACO_ERR_STATE, "postnet workspace too small: %zu < %zu", workspace_bytes,
                 ws.total);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int n = B * S;
    const float* src = d_mel_coarse;
    int c_in = N_MELS;
    float* bufs[2] = {ws.a, ws.b};
    for (int i = 0; i < POST_CONVS; ++i) {
        const bool last = i == POST_CONVS - 1;
        const int c_out = last ? N_MELS : POST_DIM;
        float* dst = last ? d_mel_post : bufs[i & 1];
        ChannelMajorOut opt = cm;                       // every layer zeroes the rows behind an utterance's own length ...
        if (!last) opt.y = nullptr;                     // ... and the last one also writes the vocoder layout
        TACO_TRY(dense_layer(h, src, c_in, h->post_conv_w[i], h->post_conv_wtc[i], h->post_conv_scale[i], h->post_conv_shift[i],
                                      last ? d_mel_coarse : nullptr, N_MELS, dst, c_out, n, S, c_in, c_out, POST_KERNEL,
                                      last ? ACT_NONE : ACT_TANH, false, st, opt));
        src = dst;
        c_in = c_out;
    }
    return B200TACO_OK;
}

B200TACO_API int32_t b200taco_postnet_fwd(b200taco_handle* h, const float* d_mel_coarse, int32_t B, int32_t S, float* d_mel_post,
                             void* d_workspace, size_t workspace_bytes, void* stream) {
    return postnet_run(h, d_mel_coarse, B, S, d_mel_post, ChannelMajorOut(), d_workspace, workspace_bytes, stream, "postnet_fwd");
}

B200TACO_API int32_t b200taco_postnet_fwd_vocoder(b200taco_handle* h, const float* d_mel_coarse, int32_t B, int32_t S,
                                                  const int32_t* d_lengths, int32_t S_out, float* d_mel_post, float* d_mel_vocoder,
                                                  void* d_workspace, size_t workspace_bytes, void* stream) {
    TACO_REQUIRE(d_mel_vocoder == nullptr || (S_out >= 1 && S_out <= S), B200TACO_ERR_INVALID,
                 "postnet_fwd_vocoder: S_out %d outside [1, %d]", S_out, S);
    TACO_REQUIRE(d_mel_vocoder != nullptr || d_lengths != nullptr, B200TACO_ERR_INVALID, "postnet_fwd_vocoder: nothing to do beyond postnet_fwd");
    ChannelMajorOut cm;
    cm.y = d_mel_vocoder; cm.lengths = d_lengths; cm.ld = S_out; cm.zero_tail = d_lengths != nullptr;
    return postnet_run(h, d_mel_coarse, B, S, d_mel_post, cm, d_workspace, workspace_bytes, stream, "postnet_fwd_vocoder");
}

B200TACO_API int32_t b200taco_measure_l2_read(void* d_buf, size_t bytes, int32_t iters, float* gbps, void* stream) {
    return measure_l2_read(d_buf, bytes, iters, gbps, static_cast<cudaStream_t>(stream));
}

B200TACO_API int32_t b200taco_measure_stream(void* d_buf, size_t per_cta, int32_t ctas, int32_t iters, int32_t mode,
                                             int32_t chunk, int32_t depth, float* gbps_per_sm, void* stream) {
    return measure_stream_probe(d_buf, per_cta, ctas, iters, mode, chunk, depth, gbps_per_sm, static_cast<cudaStream_t>(stream));
}

B200TACO_API int32_t b200taco_measure_sync_costs(int64_t* out8, void* stream) {
    TACO_REQUIRE(out8 != nullptr, B200TACO_ERR_INVALID, "measure_sync_costs: null output");
    return measure_sync_costs(reinterpret_cast<long long*>(out8), static_cast<cudaStream_t>(stream));
}

B200TACO_API int32_t b200taco_decoder_profile(b200taco_handle* h, int64_t* out, int32_t n, void* stream) {
    TACO_TRY(check_handle(h, false));
    TACO_REQUIRE(out != nullptr && n >= 5 && n <= 21, B200TACO_ERR_INVALID, "decoder_profile: bad output (n=%d)", n);
    TACO_REQUIRE(h->last_profile != nullptr, B200TACO_ERR_STATE, "decoder_profile: no decoder launch yet");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TACO_CUDA(cudaMemcpyAsync(out, h->last_profile, n * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    TACO_CUDA(cudaStreamSynchronize(st));
    return B200TACO_OK;
}

B200TACO_API int32_t b200taco_decoder_trace(b200taco_handle* h, int64_t* out, int32_t n, void* stream) {
    TACO_TRY(check_handle(h, false));
    TACO_REQUIRE(out != nullptr && n >= 1 && n <= 160 * 64 + 1024, B200TACO_ERR_INVALID, "decoder_trace: bad output (n=%d)", n);
    TACO_REQUIRE(h->last_trace != nullptr, B200TACO_ERR_STATE, "decoder_trace: the last decoder launch recorded no trace (TACO_RES_TRACE)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TACO_CUDA(cudaMemcpyAsync(out, h->last_trace, n * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    TACO_CUDA(cudaStreamSynchronize(st));
    return B200TACO_OK;
}

B200TACO_API int64_t b200taco_launch_count(int32_t reset) {
    const int64_t v = g_launches;
    if (reset) g_launches = 0;
    return v;
}

}  // extern "C"
