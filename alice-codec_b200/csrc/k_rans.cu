// k_rans.cu — FrequencyTable construction and the rANS coder, one warp per independent
// (chunk, channel) stream so that every stream is byte-identical to the reference.
//
// Replaces (reference file:line):
//   FrequencyTable::from_histogram / uniform      src/rans.rs:102-189
//   RansEncoder::{new, encode, encode_symbols, finish}   src/rans.rs:249-308
//   RansDecoder::{new, init_state, decode, decode_n}     src/rans.rs:330-381
//
// rANS is a serial recurrence on a 32-bit state; .alc carries no side information, so a
// stream cannot be split bit-exactly.  The design therefore (a) keeps ONLY the state
// recurrence on the critical path — table entries are expanded per symbol into
// {renorm limit, exact reciprocal, 4096-freq, cum} so that x/freq is one widening multiply
// and one shift, symbols are fetched 512 at a time by the whole warp and broadcast by
// shuffle, the decoder pre-loads its next stream byte — and (b) runs every stream of every
// in-flight chunk in one launch (grid = number of streams) so all SMs carry lanes.
//
// Every thread of the warp executes the same recurrence on the same values (the warp is a
// scalar engine with 32-wide fetch); only lane 0 stores.
#include "kernels.h"

namespace alice {

constexpr uint32_t kRansL = 1u << 23;   // rans.rs:244
constexpr uint32_t kProbBits = 12;      // rans.rs:50
constexpr uint32_t kProbScale = 1u << 12;

// ------------------------------------------------------------------------ table building
ALICE_D uint32_t pack_dec(uint32_t sym, uint32_t freq, uint32_t bias) {
    // sym (8) | freq-1 (12) | slot-cum (12); freq outside [1,4096] is resolved through DecAux
    uint32_t fm1 = (freq >= 1 && freq <= kProbScale) ? freq - 1 : 0;
    return sym | (fm1 << 8) | (bias << 20);
}

__global__ void ALICE_LAUNCH_BOUNDS(256, 1)
k_build_tables(const unsigned *__restrict__ hist, int n_symbols, EncSym *__restrict__ enc,
               uint32_t *__restrict__ dec_lut, DecAux *__restrict__ aux, uint16_t *__restrict__ freq_out,
               uint16_t *__restrict__ cum_out, uint8_t *__restrict__ lut8_out) {
    __shared__ uint32_t s_freq[256];
    __shared__ uint32_t s_cum[256];
    const int stream = blockIdx.x;
    const int tid = threadIdx.x;
    const unsigned *h = hist + (size_t)stream * 256;
    const int n = n_symbols;

    // The normalisation is a serial running sum in the reference (rans.rs:113-125); 256 entries
    // are cheap enough to do exactly that on one thread.
    if (tid == 0) {
        unsigned long long total = 0;
        for (int i = 0; i < n; i++) total += h[i];
        if (total == 0) {
            // FrequencyTable::uniform (rans.rs:158-189), u16 arithmetic
            uint16_t fps = (uint16_t)(kProbScale / (uint32_t)n);
            uint16_t cum = 0;
            for (int i = 0; i < n; i++) { s_cum[i] = cum; s_freq[i] = fps; cum = (uint16_t)(cum + fps); }
            s_freq[n - 1] = (uint16_t)((uint16_t)kProbScale - (uint16_t)s_cum[n - 1]);
        } else {
            uint32_t cum = 0, norm = 0;
            for (int i = 0; i < n; i++) {
                uint32_t c = h[i];
                uint32_t fr = 1;
                if (c != 0) {
                    unsigned long long t = ((unsigned long long)c * kProbScale) / total;
                    fr = t < 1 ? 1u : (uint32_t)t;
                }
                norm += fr;
                s_cum[i] = (uint16_t)cum;
                s_freq[i] = (uint16_t)fr;
                cum += fr;
            }
            if (norm != kProbScale) {
                int diff = (int)kProbScale - (int)norm;
                s_freq[n - 1] = (uint16_t)((int)s_freq[n - 1] + diff);
            }
        }
        for (int i = n; i < 256; i++) { s_freq[i] = 0; s_cum[i] = 0; }
        const uint32_t fl = s_freq[n - 1];
        DecAux a;
        a.wide_sym = (fl >= 1 && fl <= kProbScale) ? 0xffffffffu : (uint32_t)(n - 1);
        a.wide_freq = fl;
        a.pad[0] = a.pad[1] = 0;
        aux[stream] = a;
    }
    __syncthreads();

    {   // encoder entry of symbol `tid`
        const uint32_t f = s_freq[tid], cum = s_cum[tid];
        EncSym e;
        const bool zero = (f == 0), slow = (f > kProbScale);
        e.x_lim = zero ? 0u : (f >= 8192u ? 0xffffffffu : (f << 19) - 1u);
        uint32_t sh = 0, rcp = 0;
        if (!zero && !slow) {
            const uint32_t L = (f <= 1) ? 0u : (32u - (uint32_t)__clz((int)(f - 1)));  // ceil(log2 f)
            sh = L > 0 ? L - 1 : 0;
            const unsigned long long p = 1ull << (32 + sh);
            rcp = (uint32_t)((p + f - 1) / f - 1);
        }
        e.rcp = rcp;
        e.cmpl = kProbScale - f;
        e.packed = cum | (sh << 16) | ((slow ? 1u : 0u) << 24) | ((zero ? 1u : 0u) << 25);
        enc[(size_t)stream * 256 + tid] = e;
        if (freq_out) freq_out[(size_t)stream * 256 + tid] = (uint16_t)f;
        if (cum_out) cum_out[(size_t)stream * 256 + tid] = (uint16_t)cum;
    }

    // decoder LUT (rans.rs:134-144): zero-initialised, then each symbol fills [cum, min(cum+f, 4096))
    uint32_t *lut = dec_lut + (size_t)stream * kDecLutEntries;
    uint8_t *l8 = lut8_out ? lut8_out + (size_t)stream * kDecLutEntries : nullptr;
    for (uint32_t slot = tid; slot < kProbScale; slot += blockDim.x) {
        lut[slot] = pack_dec(0, s_freq[0], slot - s_cum[0]) & (n > 0 ? 0xffffffffu : 0u);
        if (l8) l8[slot] = 0;
    }
    __syncthreads();
    if (tid < n) {
        const uint32_t f = s_freq[tid], start = s_cum[tid];
        uint32_t end = start + f;
        if (end > kProbScale) end = kProbScale;
        for (uint32_t slot = start; slot < end; slot++) {
            lut[slot] = pack_dec((uint32_t)tid, f, slot - start);
            if (l8) l8[slot] = (uint8_t)tid;
        }
    }
}

void build_tables(const unsigned *d_hist, int n_streams, int n_symbols, EncSym *d_enc, uint32_t *d_dec_lut,
                  DecAux *d_aux, uint16_t *d_freq, uint16_t *d_cum, uint8_t *d_lut8, cudaStream_t st) {
    if (n_streams <= 0) return;
    ALICE_LAUNCH(k_build_tables, dim3(n_streams), dim3(256), 0, st, d_hist, n_symbols, d_enc, d_dec_lut, d_aux,
                 d_freq, d_cum, d_lut8);
}

// --------------------------------------------------------------------------------- encode
struct EncState {
    uint32_t x;
    uint8_t *wp;        // next byte goes to *--wp
    uint8_t *floor;     // lowest address that may be written
    uint32_t status;    // 0 ok, 1 overflow, 2 zero-frequency symbol
};

ALICE_D void enc_put(EncState &s, uint32_t byte, int lane) {
    if (s.wp > s.floor) {
        --s.wp;
        if (lane == 0) *s.wp = (uint8_t)byte;
    } else {
        s.status |= 1u;
    }
}

// rans.rs:269-285 RansEncoder::encode for one expanded table entry
ALICE_D void enc_symbol(EncState &s, const EncSym &e, int lane) {
    uint32_t x = s.x;
    if (x > e.x_lim) {                      // while (x >= x_max): at most two rounds for freq >= 1
        enc_put(s, x & 0xff, lane);
        x >>= 8;
        if (x > e.x_lim) {
            enc_put(s, x & 0xff, lane);
            x >>= 8;
        }
    }
    uint32_t q;
    if (e.packed >> 24) {                   // rare: freq > 4096 (malformed last symbol) or freq == 0
        if (e.packed & (1u << 25)) { s.status |= 2u; q = 0; }
        else q = x / (kProbScale - e.cmpl);
    } else {
        // x / freq for x < freq * 2^19: exact with rcp = ceil(2^(32+sh)/freq) - 1
        unsigned long long p = (unsigned long long)x * e.rcp + e.rcp;
        q = (uint32_t)(p >> 32) >> ((e.packed >> 16) & 0xff);
    }
    // (q << 12) + (x - q*freq) + cum  ==  x + cum + q * (4096 - freq)   (mod 2^32)
    s.x = x + (e.packed & 0xffffu) + q * e.cmpl;
}

__global__ void ALICE_LAUNCH_BOUNDS(32, 1)
k_rans_encode(const RansEncJob *__restrict__ jobs, const EncSym *__restrict__ enc_all,
              unsigned long long *__restrict__ results) {
    __shared__ EncSym tab[256];
    const int stream = blockIdx.x;
    const int lane = threadIdx.x;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(enc_all + (size_t)stream * 256);
        uint4 *dst = reinterpret_cast<uint4 *>(tab);
        for (int i = lane; i < 256; i += 32) dst[i] = src[i];
    }
    __syncwarp();
    const RansEncJob job = jobs[stream];
    const uint8_t *sym = job.symbols;
    EncState s;
    s.x = kRansL;
    s.wp = job.out + job.cap;
    s.floor = job.out;
    s.status = 0;

    long long i = (long long)job.n;  // symbols [0, i) remain; encode from the back (rans.rs:288-294)
    // ragged tail until the read pointer is 16-byte aligned
    while (i > 0 && ((reinterpret_cast<uintptr_t>(sym + i)) & 15) != 0) {
        enc_symbol(s, tab[__ldg(sym + i - 1)], lane);
        i--;
    }
    // full 512-symbol blocks: one coalesced 16-byte load per lane, broadcast chunk by chunk
    uint4 cur = make_uint4(0, 0, 0, 0);
    if (i >= 512) cur = __ldg(reinterpret_cast<const uint4 *>(sym + i - 512) + lane);
    while (i >= 512 && s.status == 0) {
        uint4 nxt = make_uint4(0, 0, 0, 0);
        if (i >= 1024) nxt = __ldg(reinterpret_cast<const uint4 *>(sym + i - 1024) + lane);
        for (int c = 31; c >= 0; c--) {
            uint32_t wd[4];
            wd[0] = __shfl_sync(kFullMask, cur.x, c);
            wd[1] = __shfl_sync(kFullMask, cur.y, c);
            wd[2] = __shfl_sync(kFullMask, cur.z, c);
            wd[3] = __shfl_sync(kFullMask, cur.w, c);
            EncSym e[16];  // expand all 16 entries first: keeps shared loads off the state chain
#pragma unroll
            for (int b = 0; b < 16; b++) e[b] = tab[(wd[b >> 2] >> (8 * (b & 3))) & 0xff];
#pragma unroll
            for (int b = 15; b >= 0; b--) enc_symbol(s, e[b], lane);
        }
        cur = nxt;
        i -= 512;
    }
    while (i > 0 && s.status == 0) {
        enc_symbol(s, tab[__ldg(sym + i - 1)], lane);
        i--;
    }
    // finish (rans.rs:298-308): 4 state bytes, low byte first, then the whole vector is reversed
    enc_put(s, s.x & 0xff, lane);
    enc_put(s, (s.x >> 8) & 0xff, lane);
    enc_put(s, (s.x >> 16) & 0xff, lane);
    enc_put(s, (s.x >> 24) & 0xff, lane);
    if (lane == 0) {
        results[2 * stream] = (unsigned long long)((job.out + job.cap) - s.wp);
        results[2 * stream + 1] = s.status;
    }
}

void rans_encode(const RansEncJob *d_jobs, const EncSym *d_enc, const unsigned *, unsigned long long *d_results,
                 int n_streams, cudaStream_t st) {
    if (n_streams <= 0) return;
    ALICE_LAUNCH(k_rans_encode, dim3(n_streams), dim3(32), 0, st, d_jobs, d_enc, d_results);
}

// --------------------------------------------------------------------------------- decode
struct DecState {
    uint32_t x;
    unsigned long long pos, len;
    const uint8_t *in;
    uint32_t nb;  // in[pos], pre-loaded (0 past the end)
};

// rans.rs:351-371 RansDecoder::decode
ALICE_D uint32_t dec_symbol(DecState &s, const uint32_t *lut, uint32_t f0c, uint32_t f0, uint32_t wide_sym,
                            uint32_t wide_freq) {
    const uint32_t slot = s.x & (kProbScale - 1);
    uint32_t sym;
    if (slot < f0c) {          // symbol 0 owns slots [0, freq0): no table access on the hot path
        sym = 0;
        s.x = f0 * (s.x >> kProbBits) + slot;
    } else {
        const uint32_t ent = lut[slot];
        sym = ent & 0xff;
        uint32_t f = ((ent >> 8) & 0xfff) + 1;
        if (sym == wide_sym) f = wide_freq;
        s.x = f * (s.x >> kProbBits) + (ent >> 20);   // low 32 bits of the reference's u64 expression
    }
    while (s.x < kRansL && s.pos < s.len) {
        s.x = (s.x << 8) | s.nb;
        s.pos++;
        s.nb = s.pos < s.len ? (uint32_t)__ldg(s.in + s.pos) : 0u;
    }
    return sym;
}

__global__ void ALICE_LAUNCH_BOUNDS(32, 1)
k_rans_decode(const RansDecJob *__restrict__ jobs, const uint32_t *__restrict__ lut_all,
              const DecAux *__restrict__ aux_all) {
    __shared__ uint32_t lut[kDecLutEntries];
    const int stream = blockIdx.x;
    const int lane = threadIdx.x;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(lut_all + (size_t)stream * kDecLutEntries);
        uint4 *dst = reinterpret_cast<uint4 *>(lut);
        for (int i = lane; i < kDecLutEntries / 4; i += 32) dst[i] = src[i];
    }
    __syncwarp();
    const RansDecJob job = jobs[stream];
    const DecAux aux = aux_all[stream];
    DecState s;
    s.in = job.in;
    s.len = job.len;
    s.x = 0;
    s.pos = 0;
    if (s.len >= 4) {  // rans.rs:341-347 big-endian initial state
        s.x = ((uint32_t)__ldg(s.in) << 24) | ((uint32_t)__ldg(s.in + 1) << 16) | ((uint32_t)__ldg(s.in + 2) << 8) |
              (uint32_t)__ldg(s.in + 3);
        s.pos = 4;
    }
    s.nb = s.pos < s.len ? (uint32_t)__ldg(s.in + s.pos) : 0u;
    const uint32_t ent0 = lut[0];
    uint32_t f0 = ((ent0 >> 8) & 0xfff) + 1, f0c = 0;
    if ((ent0 & 0xff) == 0 && (ent0 >> 20) == 0 && aux.wide_sym != 0) f0c = f0;  // slots [0,f0) decode to symbol 0

    uint8_t *out = job.symbols;
    unsigned long long i = 0;
    const unsigned long long n = job.n;
    while (i < n && ((reinterpret_cast<uintptr_t>(out + i)) & 3) != 0) {
        uint32_t sy = dec_symbol(s, lut, f0c, f0, aux.wide_sym, aux.wide_freq);
        if (lane == 0) out[i] = (uint8_t)sy;
        i++;
    }
    for (; i + 4 <= n; i += 4) {
        uint32_t pk = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) pk |= dec_symbol(s, lut, f0c, f0, aux.wide_sym, aux.wide_freq) << (8 * b);
        if (lane == 0) *reinterpret_cast<uint32_t *>(out + i) = pk;
    }
    for (; i < n; i++) {
        uint32_t sy = dec_symbol(s, lut, f0c, f0, aux.wide_sym, aux.wide_freq);
        if (lane == 0) out[i] = (uint8_t)sy;
    }
}

void rans_decode(const RansDecJob *d_jobs, const uint32_t *d_dec_lut, const DecAux *d_aux, int n_streams,
                 cudaStream_t st) {
    if (n_streams <= 0) return;
    ALICE_LAUNCH(k_rans_decode, dim3(n_streams), dim3(32), 0, st, d_jobs, d_dec_lut, d_aux);
}

}  // namespace alice
