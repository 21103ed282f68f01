// Compile-time specialisation of the Tacotron 2 hot path (reference src/config.py:26-59).
// b200taco_create() rejects any hparams block that differs from these.
#pragma once

namespace b200taco {

constexpr int N_SYMBOLS = 72;        // len(config.SYMBOLS), src/config.py:14-23
constexpr int EMB = 512;             // symbols_embedding_dim == encoder_embedding_dim
constexpr int ENC_CONVS = 3;
constexpr int ENC_KERNEL = 5;
constexpr int ENC_H = 256;           // BiLSTM hidden size per direction
constexpr int N_MELS = 80;
constexpr int RNN = 1024;            // decoder_rnn_dim == attention_rnn_dim
constexpr int PRENET = 256;
constexpr int ATT_DIM = 128;
constexpr int LOC_F = 32;
constexpr int LOC_K = 31;
constexpr int POST_DIM = 512;
constexpr int POST_KERNEL = 5;
constexpr int POST_CONVS = 5;

constexpr int MAX_BATCH = 64;        // rows one persistent-decoder launch carries
constexpr int MAX_TEXT_LEN = 1024;   // encoder positions per utterance

// ---- persistent decoder geometry -------------------------------------------------
constexpr int GEMM_CTAS = 128;       // CTAs that own LSTM gate rows
constexpr int UNITS_PER_CTA = RNN / GEMM_CTAS;     // 8 hidden units ...
constexpr int ROWS_PER_CTA = 4 * UNITS_PER_CTA;    // ... = 32 gate rows (i,f,g,o of each unit)
constexpr int MIN_CTAS = GEMM_CTAS;  // every CTA beyond these only takes attention / projection work items

// Activation exchange vector ("act"), one column per utterance:
//   [ prenet out | context | attention h | decoder h ]
constexpr int ACT_P = 0;
constexpr int ACT_CTX = PRENET;                    // 256
constexpr int ACT_AH = ACT_CTX + EMB;              // 768
constexpr int ACT_DH = ACT_AH + RNN;               // 1792
constexpr int ACT_ROWS = ACT_DH + RNN;             // 2816
constexpr int K_ATTN = ACT_DH;                     // attention LSTM reads rows [0, 1792)
constexpr int K_DEC = ACT_ROWS - ACT_CTX;          // decoder LSTM reads rows [256, 2816)

// Tensor-core decoder path (split-K groups of 4 CTAs): local weight chunk j of group member m -> 64-row chunk of
// the act vector.  Every availability class (prenet / context / attention h / decoder h rows) is dealt evenly
// to the four members, and the chunks of one class are contiguous in j (decoder.cu, decoder_tc_kernel).
__host__ __device__ constexpr int tc_attn_act_chunk(int m, int j) {      // 7 chunks: 1 prenet, 2 ctx, 4 ah
    return j < 1 ? m : (j < 3 ? 4 + 2 * m + (j - 1) : 12 + 4 * m + (j - 3));
}
__host__ __device__ constexpr int tc_dec_act_chunk(int m, int j) {       // 10 chunks: 2 ctx, 4 ah, 4 dh
    return j < 2 ? 4 + 2 * m + j : (j < 6 ? 12 + 4 * m + (j - 2) : 28 + 4 * m + (j - 6));
}

constexpr int PROJ_OUT = N_MELS + 1;               // 80 mel rows + the gate row
constexpr int PROJ_PAD = 96;                       // padded row length of projection partials

}  // namespace b200taco
