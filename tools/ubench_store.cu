// micro-benchmark: HBM WRITE ceilings for the store patterns of the Jacobian epilogue (dev tool).
// The Jacobian of BASELINE config 2 is 273340 rows x 1360 doubles (2.976 GB); k_accum_trie_d16 writes it as 2 KB
// row segments (one 16x16 gate block of one outcome), 8 x 64 B per warp instruction.  This program measures what
// the memory system gives for (a) cudaMemset, (b) a fully coalesced streaming store, (c) 2 KB chunks in
// sequential / scattered order, (d) the epilogue's own lane mapping, with st.cs / default / st.cg.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum { ST_CS = 0, ST_DEF = 1, ST_CG = 2 };
template <int MODE> __device__ __forceinline__ void st2(double* p, double a, double b) {
    if (MODE == ST_CS) __stcs(reinterpret_cast<double2*>(p), make_double2(a, b));
    else if (MODE == ST_CG) __stcg(reinterpret_cast<double2*>(p), make_double2(a, b));
    else *reinterpret_cast<double2*>(p) = make_double2(a, b);
}

// fully coalesced grid-stride stream: one warp instruction = 512 contiguous bytes
template <int MODE> __global__ void k_stream(double* p, size_t n2) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x)
        st2<MODE>(p + 2 * i, 1.0, 2.0);
}

// 2 KB chunks: warp w writes chunk perm(k) for k = w, w + nwarps, ...; perm(k) = (k * mult) % nchunks (mult coprime)
// LANEMAP 0: 4 instructions of 512 contiguous bytes;  1: the epilogue's mapping (row i = 8h + lane/4, cols 8z + 2(lane%4):
//            8 segments of 64 B at a 128 B stride per instruction)
template <int MODE, int LANEMAP> __global__ void k_chunks(double* p, uint32_t nchunks, uint32_t mult) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t k = w; k < nchunks; k += nw) {
        const uint32_t c = (uint32_t)(((uint64_t)k * mult) % nchunks);
        double* base = p + (size_t)c * 256;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const uint32_t off = LANEMAP ? ((8 * (t >> 1) + (lane >> 2)) * 16 + 8 * (t & 1) + 2 * (lane & 3)) : (t * 64 + 2 * lane);
            st2<MODE>(base + off, 1.0, (double)t);
        }
    }
}

// epilogue shape: a unit = (circuit c, gate g): 4 outcome rows (consecutive J rows, 1360 doubles apart) x 256 columns at
// column 256 g; units visited in order perm(u)
template <int MODE> __global__ void k_units(double* p, uint32_t ncirc, uint32_t ld, uint32_t mult) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const uint32_t nunits = ncirc * 5;
    for (uint32_t k = w; k < nunits; k += nw) {
        const uint32_t u = (uint32_t)(((uint64_t)k * mult) % nunits);
        const uint32_t c = u / 5, g = u - 5 * c;
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            double* base = p + (size_t)(4 * c + o) * ld + 256 * g;
#pragma unroll
            for (int t = 0; t < 4; ++t)
                st2<MODE>(base + (8 * (t >> 1) + (lane >> 2)) * 16 + 8 * (t & 1) + 2 * (lane & 3), 1.0, (double)t);
        }
    }
}

template <class F> static float timeit(F f, int reps = 5) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    const uint32_t ncirc = 68335, ld = 1360;
    const size_t n = (size_t)ncirc * 4 * ld;           // doubles
    const size_t bytes = n * 8;
    double* p; if (cudaMalloc(&p, bytes + 4096) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    const uint32_t nchunks = (uint32_t)(n / 256);
    auto rep = [&](const char* name, float ms, double b) { printf("%-64s %.3f ms  %.0f GB/s\n", name, ms, b / ms / 1e6); };
    rep("cudaMemsetAsync", timeit([&] { cudaMemsetAsync(p, 0, bytes); }), (double)bytes);
    for (int bps : {2, 4, 8}) {
        const int grid = 148 * bps;
        char nm[128];
        snprintf(nm, sizeof nm, "stream st.cs   (%d CTAs/SM x 256 thr)", bps); rep(nm, timeit([&] { k_stream<ST_CS><<<grid, 256>>>(p, n / 2); }), (double)bytes);
        snprintf(nm, sizeof nm, "stream default (%d CTAs/SM x 256 thr)", bps); rep(nm, timeit([&] { k_stream<ST_DEF><<<grid, 256>>>(p, n / 2); }), (double)bytes);
    }
    const int grid = 148 * 2;   // 16 warps/SM as the accumulate kernel
    const double cb = (double)nchunks * 2048;
    rep("2KB chunks sequential, contiguous lanes, st.cs", timeit([&] { k_chunks<ST_CS, 0><<<grid, 256>>>(p, nchunks, 1u); }), cb);
    rep("2KB chunks scattered,  contiguous lanes, st.cs", timeit([&] { k_chunks<ST_CS, 0><<<grid, 256>>>(p, nchunks, 2654435761u % nchunks | 1u); }), cb);
    rep("2KB chunks sequential, epilogue lanes,   st.cs", timeit([&] { k_chunks<ST_CS, 1><<<grid, 256>>>(p, nchunks, 1u); }), cb);
    rep("2KB chunks scattered,  epilogue lanes,   st.cs", timeit([&] { k_chunks<ST_CS, 1><<<grid, 256>>>(p, nchunks, 2654435761u % nchunks | 1u); }), cb);
    rep("2KB chunks scattered,  epilogue lanes,   default", timeit([&] { k_chunks<ST_DEF, 1><<<grid, 256>>>(p, nchunks, 2654435761u % nchunks | 1u); }), cb);
    rep("2KB chunks scattered,  epilogue lanes,   st.cg", timeit([&] { k_chunks<ST_CG, 1><<<grid, 256>>>(p, nchunks, 2654435761u % nchunks | 1u); }), cb);
    const double ub = (double)ncirc * 5 * 4 * 2048;
    // mult must be coprime to nunits = 341675 = 5^2 * 79 * 173
    rep("units (4 rows x 2KB) sequential, st.cs", timeit([&] { k_units<ST_CS><<<grid, 256>>>(p, ncirc, ld, 1u); }), ub);
    rep("units (4 rows x 2KB) scattered,  st.cs", timeit([&] { k_units<ST_CS><<<grid, 256>>>(p, ncirc, ld, 100003u); }), ub);
    rep("units (4 rows x 2KB) scattered,  default", timeit([&] { k_units<ST_DEF><<<grid, 256>>>(p, ncirc, ld, 100003u); }), ub);
    rep("units scattered, st.cs, 4 CTAs/SM (32 warps)", timeit([&] { k_units<ST_CS><<<148 * 4, 256>>>(p, ncirc, ld, 100003u); }), ub);
    rep("units scattered, st.cs, 1 CTA/SM (8 warps)", timeit([&] { k_units<ST_CS><<<148, 256>>>(p, ncirc, ld, 100003u); }), ub);
    printf("cuda status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
