// red_span.cu -- microbenchmark (development aid): RED.64 throughput against the ADDRESS SPAN of
// a fixed hot set. NH hot cells (8-byte counters), cell j placed inside its own window of S cells
// at a hashed offset, so the hot set always occupies NH distinct sectors (S >= 4) while the span
// grows from NH*8 bytes to NH*S*8 bytes. If throughput falls with S although the hot set stays
// L2-resident, the cost is in address translation / slice mapping, not in L2 capacity or HBM.
// Also: the sierpinski_3d@512^3 address pattern itself, generated arithmetically.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_span red_span.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

// mode 0: hashed windows; mode 1: sierpinski tetrahedron cells of a 512^3 grid (9 levels, one of 4
// corner moves per level); mode 2: the same cells, z-major compacted rank (dense 2 MiB hot set)
__global__ void red_kernel(u64 *buf, int mode, uint32_t nh_mask, uint32_t S, int iters, uint32_t seed)
{
    uint32_t s = mix32(seed + blockIdx.x*blockDim.x + threadIdx.x);
    for (int i = 0; i < iters; ++i)
    {
        s = s*1664525u + 1013904223u;
        const uint32_t r = mix32(s);
        u64 pos;
        if (mode == 0)
        {
            const uint32_t j = r & nh_mask;
            pos = (u64)j*S + (mix32(j*2654435761u) & (S - 1));
        }
        else
        {
            uint32_t x = 0, y = 0, z = 0, rank = 0;
            #pragma unroll
            for (int l = 0; l < 9; ++l)
            {
                const uint32_t c = (r >> (2*l)) & 3u;
                x |= (c == 1u) << l; y |= (c == 2u) << l; z |= (c == 3u) << l;
                rank |= c << (2*l);
            }
            pos = mode == 1 ? (u64)x + 512ull*y + 262144ull*z : (u64)rank;
        }
        atomicAdd(&buf[pos], 1ull);
    }
}

// mode 3: sierpinski_3d cells through a ROW DIRECTORY: dir[row] (row = cell >> 9, one 4 KiB row of
// 512 cells) holds the row's slot in a compact tile, allocated on first touch; the RED goes to
// tile[slot*512 + (cell & 511)]. Span of the tile = (touched rows) * 4 KiB.
__global__ void red_dir_kernel(u64 *tile, unsigned int *dir, unsigned int *next_slot, int iters, uint32_t seed, int cached)
{
    uint32_t s = mix32(seed + blockIdx.x*blockDim.x + threadIdx.x);
    for (int i = 0; i < iters; ++i)
    {
        s = s*1664525u + 1013904223u;
        const uint32_t r = mix32(s);
        uint32_t x = 0, y = 0, z = 0;
        #pragma unroll
        for (int l = 0; l < 9; ++l)
        {
            const uint32_t c = (r >> (2*l)) & 3u;
            x |= (c == 1u) << l; y |= (c == 2u) << l; z |= (c == 3u) << l;
        }
        const uint32_t row = y + 512u*z;
        unsigned int slot = cached ? __ldg(&dir[row]) : __ldcg(&dir[row]);
        if (slot == 0xffffffffu)
        {
            slot = __ldcg(&dir[row]);
            if (slot == 0xffffffffu)
            {
                const unsigned int mine = atomicAdd(next_slot, 1u);
                const unsigned int old = atomicCAS(&dir[row], 0xffffffffu, mine);
                slot = old == 0xffffffffu ? mine : old;     /* a lost race leaks one slot: fine here */
            }
        }
        atomicAdd(&tile[(u64)slot*512u + x], 1ull);
    }
}

int main()
{
    const uint32_t NH = 1u << 18;
    u64 *buf;
    const size_t bytes = (size_t)1 << 30;
    CK(cudaMalloc(&buf, bytes));
    CK(cudaMemset(buf, 0, bytes));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int grid = 148*3, tpb = 256, iters = 4096;
    auto run = [&](const char *name, int mode, uint32_t S) {
        red_kernel<<<grid,tpb>>>(buf, mode, NH - 1, S, 256, 1);   // warm
        CK(cudaEventRecord(e0));
        red_kernel<<<grid,tpb>>>(buf, mode, NH - 1, S, iters, 7);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("%-34s span %8.1f MiB  %.3e RED/s\n", name, mode == 0 ? NH*(double)S*8/1048576.0 : (mode == 1 ? 1024.0 : 2.0),
               (double)grid*tpb*iters/(ms*1e-3));
    };
    for (uint32_t S = 1; S <= 512; S *= 2) run("hashed windows, 2^18 hot cells", 0, S);
    run("sierpinski_3d cells, x+512y+512^2z", 1, 0);
    run("sierpinski_3d cells, compact rank", 2, 0);
    {
        unsigned int *dir, *next;
        CK(cudaMalloc(&dir, 262144*4)); CK(cudaMalloc(&next, 4));
        for (int cached = 0; cached < 2; ++cached)
        {
            CK(cudaMemset(dir, 0xff, 262144*4)); CK(cudaMemset(next, 0, 4));
            red_dir_kernel<<<grid,tpb>>>(buf, dir, next, 256, 1, cached);
            CK(cudaEventRecord(e0));
            red_dir_kernel<<<grid,tpb>>>(buf, dir, next, iters, 7, cached);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            unsigned int used; CK(cudaMemcpy(&used, next, 4, cudaMemcpyDeviceToHost));
            printf("%-34s span %8.1f MiB  %.3e RED/s (%u rows)\n", cached ? "sierp3d via row directory (ld.nc)" : "sierp3d via row directory (ld.cg)",
                   used*4096.0/1048576.0, (double)grid*tpb*iters/(ms*1e-3), used);
        }
    }
    return 0;
}
