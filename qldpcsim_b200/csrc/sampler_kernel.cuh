// sampler_kernel.cuh -- on-device depolarizing sampler + syndrome generator (stand-in for the reference's Stim
// circuit and sampler, simulator.py:43-160, :196-197; distribution per SURVEY.md section 8d).
//
// Per data qubit one uniform 32-bit integer x is drawn from Philox4x32-10 keyed by (seed) with counter
// (global shot index, qubit/4); the Pauli is X if x < t1, Y if t1 <= x < t2, Z if t2 <= x < t3 with
// t_k = round(k p/3 * 2^32) (the reference's PAULI_CHANNEL_1(p/3, p/3, p/3), simulator.py:107).  errX = X|Y,
// errZ = Z|Y, syn_z = Hz errX, syn_x = Hx errZ (mod 2): the record columns of simulator.py:141-144.
// Because the counter is the GLOBAL shot index the batch does not depend on how shots are sharded over GPUs.
// One warp per shot: lane <-> output word (32 qubits = 8 Philox blocks), then lane <-> check.
#pragma once
#include "common.cuh"
#include "hard_kernels.cuh"

namespace qldpc {

struct SampleArgs {
    GraphDev gz, gx;
    double p;
    unsigned long long seed;
    long long first_shot, shots;
    uint32_t *errx, *errz, *synz, *synx;
    const uint32_t *hcol_z, *hcol_x;       // [n][kColStride] bit-packed columns of Hz / Hx
};

__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t out[4])
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = (unsigned long long)M0 * c0, p1 = (unsigned long long)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// thresholds on the 32-bit draw; computed identically on host (tests) and device
__host__ __device__ inline void pauli_thresholds(double p, unsigned long long t[3])
{
    for (int k = 0; k < 3; ++k) {
        double v = (double)(k + 1) * p / 3.0 * 4294967296.0;
        unsigned long long u = (unsigned long long)(v + 0.5);
        t[k] = u > 4294967296ull ? 4294967296ull : u;
    }
}

// VH: 128-bit loads per column of H (4*VH >= words(m) of both matrices); VH = 0: row-wise parities (more than 1024 checks)
template <int VH>
__global__ void __launch_bounds__(256) sample_kernel(SampleArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nw = a.gz.nw, n = a.gz.n;
    uint32_t *ex = reinterpret_cast<uint32_t *>(smem) + (size_t)warp * 2 * nw;
    uint32_t *ez = ex + nw;
    unsigned long long thr[3];
    pauli_thresholds(a.p, thr);
    const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int nblk = (n + 3) / 4;                             // Philox blocks per shot: block b draws qubits 4b .. 4b+3
    const int nblk_pad = (nw * 8 + 31) & ~31;                 // whole rounds of 32 blocks = 4 words
    for (long long s = warp_global; s < a.shots; s += nwarps) {
        const unsigned long long gs = (unsigned long long)(a.first_shot + s);
        // ---- draws: lane <-> block (all 32 lanes busy; a word is 8 consecutive blocks = 8 consecutive lanes, whose nibbles are
        // OR-combined by three xor-shuffles)
        for (int b0 = 0; b0 < nblk_pad; b0 += 32) {
            const int b = b0 + lane;
            uint32_t nx = 0, nz = 0;
            if (b < nblk) {
                uint32_t r[4];
                philox4x32_10((uint32_t)gs, (uint32_t)(gs >> 32), (uint32_t)b, 0u, k0, k1, r);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (4 * b + q < n) {
                        const unsigned long long x = r[q];
                        const bool X = x < thr[0], Y = (x >= thr[0]) && (x < thr[1]), Z = (x >= thr[1]) && (x < thr[2]);
                        nx |= (uint32_t)(X || Y) << q;
                        nz |= (uint32_t)(Z || Y) << q;
                    }
                }
            }
            uint32_t wx = nx << (4 * (lane & 7)), wz = nz << (4 * (lane & 7));
#pragma unroll
            for (int d = 1; d < 8; d <<= 1) {
                wx |= __shfl_xor_sync(0xffffffffu, wx, d);
                wz |= __shfl_xor_sync(0xffffffffu, wz, d);
            }
            const int w = b >> 3;
            if ((lane & 7) == 0 && w < nw) {
                ex[w] = wx; ez[w] = wz;
                a.errx[s * nw + w] = wx;
                a.errz[s * nw + w] = wz;
            }
        }
        __syncwarp();
        // ---- syndromes: syn_z = Hz errX, syn_x = Hx errZ as XORs of the bit-packed columns selected by the (sparse) errors
        if (VH > 0) {
            constexpr int V = VH > 0 ? VH : 1;
            uint32_t sz[4 * V], sx[4 * V];
            xor_columns<V>(reinterpret_cast<const uint4 *>(a.hcol_z), nw, ex, nullptr, lane, sz);
            xor_columns<V>(reinterpret_cast<const uint4 *>(a.hcol_x), nw, ez, nullptr, lane, sx);
#pragma unroll
            for (int k = 0; k < 4 * V; ++k) {
                const uint32_t rz = __reduce_xor_sync(0xffffffffu, sz[k]), rx = __reduce_xor_sync(0xffffffffu, sx[k]);
                if (lane == 0 && k < a.gz.mw) a.synz[s * a.gz.mw + k] = rz;
                if (lane == 0 && k < a.gx.mw) a.synx[s * a.gx.mw + k] = rx;
            }
        } else {
            for (int w = 0; w < a.gz.mw; ++w) {                 // more than 1024 checks: row-wise parities
                const int i = w * 32 + lane;
                uint32_t par = 0;
                if (i < a.gz.m) for (int x = a.gz.row_ptr[i]; x < a.gz.row_ptr[i + 1]; ++x) par ^= get_bit(ex, a.gz.col_idx[x]);
                const uint32_t bal = __ballot_sync(0xffffffffu, par);
                if (lane == 0) a.synz[s * a.gz.mw + w] = bal;
            }
            for (int w = 0; w < a.gx.mw; ++w) {
                const int i = w * 32 + lane;
                uint32_t par = 0;
                if (i < a.gx.m) for (int x = a.gx.row_ptr[i]; x < a.gx.row_ptr[i + 1]; ++x) par ^= get_bit(ez, a.gx.col_idx[x]);
                const uint32_t bal = __ballot_sync(0xffffffffu, par);
                if (lane == 0) a.synx[s * a.gx.mw + w] = bal;
            }
        }
        __syncwarp();
    }
}

}  // namespace qldpc
