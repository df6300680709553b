/*
ffr_kernels.cuh -- the render kernel family and its helper kernels (sm_100a).

K1  render_kernel<D,RCAP,AFFINE_ONLY>   chaos game iterate + scatter
        == BufferRenderer::_render_batch (renderers/buffer_renderer.hpp:150-250) with
           RenderIterator::_init / iterate (renderers/render_iterator.hpp:52-60,106-139)
K2  add_buffer_kernel                    addBuffer (buffer_renderer.hpp:375-391), also the
                                         multi-GPU peer-memory reduce
K3  hist_sum_max_kernel                  histogramSum / histogramMax (:483-509)
T1  iterate_points_kernel, isaac_words_kernel   test hooks on the same device functions
M1  atomic_bench_kernel                  random-atomic microbenchmark (measured scatter roofline)

Execution model of K1: persistent blocks (grid = SMs x resident blocks), FFR_TPB chains per
block advanced in lock step; a block takes "chain groups" (FFR_TPB consecutive chains) from a
global work counter until none are left. Chain k is seeded with splitmix64(base_seed + k), so
the result does not depend on which block, SM or GPU runs it.
*/

#pragma once

#include "ffr_device.cuh"
#include "ffr_params.cuh"

#define FFR_SMEM_RNG_BYTES_W(WB) (FFR_RNG_WORDS*FFR_TPB*(WB))
#define FFR_SMEM_RNG_BYTES FFR_SMEM_RNG_BYTES_W(8)
#ifndef FFR_DIRECT_MINB
#define FFR_DIRECT_MINB 2
#endif

__device__ __forceinline__ void stage_blob(void *dst, const void *src, uint32_t bytes)
{
    const u64 *s = (const u64*)src;
    u64 *d = (u64*)dst;
    for (uint32_t i = threadIdx.x; i < bytes/8; i += blockDim.x)
        d[i] = s[i];
    __syncthreads();
}

/* add the 16-bit per-xform selection counters packed in pk0/pk1 to the block's counters */
__device__ __forceinline__ void flush_packed(unsigned long long *s_xf, u64 &pk0, u64 &pk1)
{
#pragma unroll
    for (int f = 0; f < 4; ++f)
    {
        const unsigned a0 = (unsigned)((pk0 >> (16*f)) & 0xffffu);
        const unsigned a1 = (unsigned)((pk1 >> (16*f)) & 0xffffu);
        if (a0) atomicAdd(&s_xf[f],(unsigned long long)a0);
        if (a1) atomicAdd(&s_xf[4+f],(unsigned long long)a1);
    }
    pk0 = pk1 = 0;
}

/* colour loops: registers (fully unrolled, predicated) for RCAP <= 4, local memory beyond */
#define FOR_COLOR(i) _Pragma("unroll") \
    for (int i = 0; i < (RCAP <= 4 ? RCAP : (int)r); ++i) if (RCAP > 4 || i < (int)r)

template <typename T, int D, int RCAP> struct ChainState
{
    RngT<T> rng;
    T p[D];
    T c[RCAP > 0 ? RCAP : 1];
};

/* cold path, kept out of line and by value: RenderIterator::init() after a bad value
   (render_iterator.hpp:52-60 via buffer_renderer.hpp:185) */
template <typename T, int D, int RCAP, bool AFFINE_ONLY>
__device__ __noinline__ ChainState<T,D,RCAP> chain_reinit(const DevFlameT<T> *fl, RngT<T> rng)
{
    ChainState<T,D,RCAP> st;
    const DevXFormT<T> *xfs = blob_xforms(fl);
    const DevVarT<T> *vars = blob_vars(fl);
    const uint32_t r = fl->r;
    T p[D];
#pragma unroll
    for (int i = 0; i < D; ++i)
        p[i] = 2.0*rng.num() - 1.0;
    for (int s = 0; s < Real<T>::settle_iters; ++s)
    {
        uint32_t xi = select_xform(fl,rng);
        xform_apply<T,D,AFFINE_ONLY>(xfs[xi],vars,rng,p,p);
    }
    FOR_COLOR(i)
        st.c[i] = rng.num();
#pragma unroll
    for (int i = 0; i < D; ++i)
        st.p[i] = p[i];
    st.rng = rng;
    return st;
}

template <typename T, int D, int RCAP, bool AFFINE_ONLY>
__global__ void __launch_bounds__(FFR_TPB,AFFINE_ONLY ? 3 : FFR_DIRECT_MINB) render_kernel(const RenderParams prm)
{
    typedef typename Real<T>::word W;   /* generator word == histogram counter type */
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ unsigned int s_group;
    W *rng_base = (W*)smem;
    DevFlameT<T> *fl = (DevFlameT<T>*)(smem + FFR_SMEM_RNG_BYTES_W(sizeof(W)));
    stage_blob(fl,prm.blob,prm.blob_bytes);

    const DevXFormT<T> *xfs = blob_xforms(fl);
    const DevVarT<T> *vars = blob_vars(fl);
    const uint32_t nx = fl->num_xforms;
    const uint32_t r = fl->r;
    const uint32_t cellsz = fl->cell;
    const bool has_final = fl->has_final;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const T *__restrict__ colors = (const T*)prm.colors;
    W *__restrict__ buffer = (W*)prm.buffer;
    const bool warp_agg = prm.scatter_mode == FFR_SCATTER_WARP_AGG;
    const bool discard = prm.scatter_mode == FFR_SCATTER_DISCARD || prm.scatter_mode == FFR_SCATTER_TRACE;
    u64 *__restrict__ trace = prm.scatter_mode == FFR_SCATTER_TRACE ? prm.trace : nullptr;

    RngT<T> rng;
    rng.bind(rng_base,tid);
    rng.a = rng.b = rng.c = 0;
    rng.cnt = 0;

    /* per-thread statistics (buffer_renderer.hpp:156-160), merged at kernel end (:232-246) */
    u64 n_iter = 0, n_plot = 0;
    u64 xfc0 = 0, xfc1 = 0;   /* lane i counts selections of xform i (and i+32) for its warp */
    /* up to 8 xforms: ++xf_dist[xf_id] as 16-bit fields packed in two registers per thread,
       flushed to shared counters every 32768 iterations (one ballot per xform otherwise) */
    __shared__ unsigned long long s_xf[8];
    const bool packed_xf = nx <= 8;
    const bool long_chain = prm.chain_len > 32768;  /* 16-bit fields: flush mid-chain too */
    u64 pk0 = 0, pk1 = 0;
    if (tid < 8)
        s_xf[tid] = 0;
    __syncthreads();
    T pmin[D], pmax[D];
#pragma unroll
    for (int i = 0; i < D; ++i)
    {
        pmin[i] = INFINITY;
        pmax[i] = -INFINITY;
    }

    const u64 num_groups = (prm.chain_count + FFR_TPB - 1) / FFR_TPB;
    const int chain_len = (int)prm.chain_len;   /* host guarantees chain_len < 2^31 */

    for (;;)
    {
        /* block-uniform: thread 0 takes the next chain group unless the bad value limit
           was hit (buffer_renderer.hpp:152-153) */
        if (tid == 0)
            s_group = (*(volatile uint32_t*)&prm.stats->abort) ? 0xffffffffu
                                                               : atomicAdd(prm.work_counter,1u);
        __syncthreads();
        const u64 g = s_group;
        __syncthreads();
        if (g >= num_groups)
            break;
        const u64 kk = g*FFR_TPB + tid;
        const bool active = kk < prm.chain_count;
        const int len = !active ? 0 :
            ((kk+1 == prm.chain_count && prm.last_len) ? (int)prm.last_len : chain_len);
        /* rng::setSeed((u64)seed_k) */
        rng.seed(splitmix64(prm.base_seed + prm.chain_first + kk));

        T p[D], pf[D];
        T c[RCAP > 0 ? RCAP : 1], cf[RCAP > 0 ? RCAP : 1];
        bool dead = false;
        /* RenderIterator::_init: p = randPoint (flame_rng.hpp:151-158) */
#pragma unroll
        for (int i = 0; i < D; ++i)
            p[i] = 2.0*rng.num() - 1.0;

        /* iterations -53..-1 are the settle iterations of _init (no stats, no plotting);
           sharing the loop keeps one inlined copy of the xform interpreter */
        for (int it = -Real<T>::settle_iters; it < chain_len; ++it)
        {
            if (it == 0 && RCAP > 0)
            {
                FOR_COLOR(i)
                    c[i] = rng.num();
            }
            const bool alive = !dead && it < len;
            uint32_t xi = 0xffffffffu;
            if (alive)
            {
                /* RenderIterator::iterate, render_iterator.hpp:106-139 */
                xi = select_xform(fl,rng);
                const DevXFormT<T> &xf = xfs[xi];
                xform_apply<T,D,AFFINE_ONLY>(xf,vars,rng,p,p);
                if (it >= 0)
                {
                    if (RCAP > 0 && (xf.flags & XF_HAS_COLOR))
                    {
                        const T s = xf.color_speed;
                        FOR_COLOR(i)
                            c[i] = (1.0-s)*c[i] + s*__ldg(colors + xf.color_off + i);
                    }
                    if (has_final)
                    {
                        const DevXFormT<T> &xff = xfs[nx];
                        xform_apply<T,D,AFFINE_ONLY>(xff,vars,rng,p,pf);
                        if (RCAP > 0)
                        {
                            if (xff.flags & XF_HAS_COLOR)
                            {
                                const T s = xff.color_speed;
                                FOR_COLOR(i)
                                    cf[i] = (1.0-s)*c[i] + s*__ldg(colors + xff.color_off + i);
                            }
                            else
                            {
                                FOR_COLOR(i)
                                    cf[i] = c[i];
                            }
                        }
                    }
                    else
                    {
#pragma unroll
                        for (int i = 0; i < D; ++i)
                            pf[i] = p[i];
                        if (RCAP > 0)
                        {
                            FOR_COLOR(i)
                                cf[i] = c[i];
                        }
                    }
                    /* _render_batch body, buffer_renderer.hpp:171-229 */
                    ++n_iter;
                    bool bad = false;
#pragma unroll
                    for (int i = 0; i < D; ++i)
                        bad |= bad_value(p[i]);
                    if (bad) /* :175-186 */
                    {
                        u64 idx = atomicAdd(&prm.stats->n_bad,1ULL);
                        if (idx < FFR_MAX_BAD_RECORDED)
                        {
                            prm.stats->bad_xf[idx] = xf.json_id;
#pragma unroll
                            for (int i = 0; i < D; ++i)
                                prm.stats->bad_pt[idx][i] = (double)p[i];
                        }
                        if (idx + 1 > prm.bv_limit)
                        {
                            prm.stats->abort = 1;
                            dead = true;
                        }
                        else
                        {
                            /* iter.init(): pf, cf keep their stale values */
                            ChainState<T,D,RCAP> st = chain_reinit<T,D,RCAP,AFFINE_ONLY>(fl,rng);
                            rng = st.rng;
#pragma unroll
                            for (int i = 0; i < D; ++i)
                                p[i] = st.p[i];
                            FOR_COLOR(i)
                                c[i] = st.c[i];
                        }
                    }
                    if (!dead)
                    {
#pragma unroll
                        for (int i = 0; i < D; ++i) /* :188-194 */
                        {
                            pmin[i] = (p[i] < pmin[i]) ? p[i] : pmin[i];
                            pmax[i] = (p[i] > pmax[i]) ? p[i] : pmax[i];
                        }
                        /* inclusive bounds on pf (render_iterator.hpp:72-79); NaN is out (Q4) */
                        bool inb = true;
#pragma unroll
                        for (int i = 0; i < D; ++i)
                            inb &= (pf[i] >= fl->lo[i]) && (pf[i] <= fl->hi[i]);
                        if (inb)
                        {
                            ++n_plot;
                            /* :202-209: truncating double -> index per dimension */
                            u64 bi = to_index((pf[0] - fl->lo[0]) * fl->mult_d[0]);
#pragma unroll
                            for (int i = 1; i < D; ++i)
                                bi += to_index((pf[i] - fl->lo[i]) * fl->mult_d[i]) * fl->mult_i[i];
                            if (trace)
                                trace[(u64)it*prm.chain_count + kk] = bi;
                            W *cell = buffer + bi*cellsz;
                            if (warp_agg)
                            {
                                /* hits of colliding lanes are merged before they leave the SM */
                                const unsigned peers = __match_any_sync(__activemask(),bi);
                                if ((int)(__ffs(peers) - 1) == lane)
                                    hist_add(cell,(unsigned)__popc(peers));
                            }
                            else if (!discard)
                                hist_add(cell,1u); /* :211-215 */
                            if (RCAP > 0 && !discard)
                            {
                                FOR_COLOR(i) /* :217-229 */
                                    atomicAdd((T*)(cell + 1 + i),cf[i]);
                            }
                        }
                    }
                }
            }
            if (it >= 0 && packed_xf)
            {
                /* xi == 0xffffffff (not alive) matches neither range */
                const u64 inc = 1ULL << ((xi & 3u)*16u);
                pk0 += (xi < 4u) ? inc : 0ULL;
                if (nx > 4)
                    pk1 += (xi - 4u < 4u) ? inc : 0ULL;
                if (long_chain && (it & 0x7fff) == 0x7fff)
                    flush_packed(s_xf,pk0,pk1);
            }
            else if (it >= 0)
            {
                /* ++xf_dist[xf_id] (:172): one ballot per xform, lane i owns xform i's counter */
                __syncwarp();
                for (uint32_t i = 0; i < nx; ++i)
                {
                    const unsigned m = __ballot_sync(0xffffffffu,xi == i);
                    if ((uint32_t)lane == (i & 31u))
                    {
                        if (i < 32) xfc0 += __popc(m);
                        else xfc1 += __popc(m);
                    }
                }
            }
        }
        if (packed_xf)
            flush_packed(s_xf,pk0,pk1);   /* fields hold at most one chain's selections */
    }

    /* merge statistics, buffer_renderer.hpp:232-246 */
    __syncthreads();
    if (packed_xf && (uint32_t)tid < nx && s_xf[tid])
        atomicAdd(&prm.stats->xf_dist[tid],s_xf[tid]);
    __syncwarp();
    if ((uint32_t)lane < nx && xfc0)
        atomicAdd(&prm.stats->xf_dist[lane],xfc0);
    if ((uint32_t)lane + 32 < nx && xfc1)
        atomicAdd(&prm.stats->xf_dist[lane+32],xfc1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        n_iter += __shfl_xor_sync(0xffffffffu,n_iter,o);
        n_plot += __shfl_xor_sync(0xffffffffu,n_plot,o);
#pragma unroll
        for (int i = 0; i < D; ++i)
        {
            T a = __shfl_xor_sync(0xffffffffu,pmin[i],o);
            T b = __shfl_xor_sync(0xffffffffu,pmax[i],o);
            pmin[i] = (a < pmin[i]) ? a : pmin[i];
            pmax[i] = (b > pmax[i]) ? b : pmax[i];
        }
    }
    if (lane == 0)
    {
        if (n_iter) atomicAdd(&prm.stats->s_iter,n_iter);
        if (n_plot) atomicAdd(&prm.stats->s_plot,n_plot);
#pragma unroll
        for (int i = 0; i < D; ++i)
        {
            atomicMin(&prm.stats->pt_min[i],f64_to_ordered((double)pmin[i]));
            atomicMax(&prm.stats->pt_max[i],f64_to_ordered((double)pmax[i]));
        }
    }
}

/* K1b: the same chain semantics with a per-iteration REGROUP of the block's chains by the
   xform each one selected, so a warp interprets one op list instead of serialising over every
   xform its 32 lanes happened to draw (measured on csci6360: 14 variation bodies per
   warp-iteration instead of 3.2). Chain state (point, colour, generator words) lives in shared
   memory indexed by chain slot; every iteration
     A  the slot's OWNER thread draws the xform index from the slot's own ISAAC stream
     S  counting sort of the 256 slots by xform index (match_any ranks + one warp of prefix sums)
     B  thread j advances slot perm[j]: xform, final xform, statistics, scatter
   Which thread advances a chain does not change its arithmetic or its stream, so results are
   identical to K1 (and to the oracle). Needs num_xforms <= 31 and color_dims <= 4. */
#define FFR_NWARPS (FFR_TPB/32)
#ifndef FFR_REGROUP_MINB
#define FFR_REGROUP_MINB 2
#endif

template <typename T, int D, int RCAP>
__global__ void __launch_bounds__(FFR_TPB,FFR_REGROUP_MINB) render_kernel_regroup(const RenderParams prm)
{
    typedef typename Real<T>::word W;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ unsigned int s_group;
    __shared__ unsigned int s_xi[FFR_TPB];
    __shared__ unsigned short s_perm[FFR_TPB];
    __shared__ unsigned int s_wcnt[FFR_NWARPS][32];
    W *rng_base = (W*)smem;                          /* randmem columns only (16 words/slot) */
    W *rsl_base = (W*)prm.rsl_scratch + (size_t)blockIdx.x*16*FFR_TPB; /* randrsl: L2-resident */
    W *st_a = rng_base + 16*FFR_TPB;                 /* randa, randb, randc, randcnt per slot */
    W *st_b = st_a + FFR_TPB;
    W *st_c = st_b + FFR_TPB;
    W *st_n = st_c + FFR_TPB;
    T *sp = (T*)(st_n + FFR_TPB);                    /* p[d][slot] */
    T *sc = sp + D*FFR_TPB;                          /* c[i][slot], RCAP rows */
    DevFlameT<T> *fl = (DevFlameT<T>*)(sc + RCAP*FFR_TPB);
    stage_blob(fl,prm.blob,prm.blob_bytes);

    const DevXFormT<T> *xfs = blob_xforms(fl);
    const DevVarT<T> *vars = blob_vars(fl);
    const uint32_t nx = fl->num_xforms;
    const uint32_t r = fl->r;
    const uint32_t cellsz = fl->cell;
    const bool has_final = fl->has_final;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const T *__restrict__ colors = (const T*)prm.colors;
    W *__restrict__ buffer = (W*)prm.buffer;
    const bool warp_agg = prm.scatter_mode == FFR_SCATTER_WARP_AGG;
    const bool discard = prm.scatter_mode == FFR_SCATTER_DISCARD || prm.scatter_mode == FFR_SCATTER_TRACE;
    u64 *__restrict__ trace = prm.scatter_mode == FFR_SCATTER_TRACE ? prm.trace : nullptr;

    u64 n_iter = 0, n_plot = 0;
    u64 xfc = 0;             /* warp 0, lane k: selections of xform k */
    T pmin[D], pmax[D];
#pragma unroll
    for (int i = 0; i < D; ++i)
    {
        pmin[i] = INFINITY;
        pmax[i] = -INFINITY;
    }

    const u64 num_groups = (prm.chain_count + FFR_TPB - 1) / FFR_TPB;
    const int chain_len = (int)prm.chain_len;   /* host guarantees chain_len < 2^31 */

#define LOAD_RNG(R,slot) do { (R).col = rng_base + (slot); (R).rcol = rsl_base + (slot); (R).a = st_a[slot]; (R).b = st_b[slot]; \
        (R).c = st_c[slot]; (R).cnt = (int)st_n[slot]; } while (0)
#define STORE_RNG(R,slot) do { st_a[slot] = (R).a; st_b[slot] = (R).b; st_c[slot] = (R).c; \
        st_n[slot] = (W)(R).cnt; } while (0)

    for (;;)
    {
        if (tid == 0)
            s_group = (*(volatile uint32_t*)&prm.stats->abort) ? 0xffffffffu
                                                               : atomicAdd(prm.work_counter,1u);
        __syncthreads();
        const u64 g = s_group;
        __syncthreads();
        if (g >= num_groups)
            break;
        const u64 kk = g*FFR_TPB + tid;
        const bool active = kk < prm.chain_count;
        const int len = !active ? 0 :
            ((kk+1 == prm.chain_count && prm.last_len) ? (int)prm.last_len : chain_len);
        bool dead = false;   /* owner-side flag of slot tid */
        {
            RngT<T> rng;
            rng.col = rng_base + tid;
            rng.rcol = rsl_base + tid;
            rng.seed(splitmix64(prm.base_seed + prm.chain_first + kk));
#pragma unroll
            for (int i = 0; i < D; ++i)
                sp[i*FFR_TPB + tid] = 2.0*rng.num() - 1.0;
            STORE_RNG(rng,tid);
        }
        s_xi[tid] = 0;

        for (int it = -Real<T>::settle_iters; it < chain_len; ++it)
        {
            /* ---- A: owner draws ---- */
            RngT<T> rng;
            LOAD_RNG(rng,tid);
            if (it == 0 && RCAP > 0)
            {
                FOR_COLOR(i)
                    sc[i*FFR_TPB + tid] = rng.num();
            }
            /* a slot whose chain hit the bad value limit is flagged 0xfffffffe by its worker */
            dead = dead || (s_xi[tid] == 0xfffffffeu);
            const bool alive = !dead && it < len;
            uint32_t key = nx;
            if (alive)
                key = select_xform(fl,rng);
            STORE_RNG(rng,tid);
            s_xi[tid] = alive ? key : 0xffffffffu;
            /* ---- S: counting sort of slots by key ---- */
            s_wcnt[warp][lane] = 0;
            __syncwarp();
            const unsigned peers = __match_any_sync(0xffffffffu,key);
            const unsigned rank = __popc(peers & ((1u << lane) - 1u));
            if (rank == 0)
                s_wcnt[warp][key] = __popc(peers);
            __syncthreads();
            {
                /* every warp computes the same prefix sums (lane = key) instead of one warp
                   computing them for all behind another barrier */
                unsigned total = 0, before = 0;
#pragma unroll
                for (int w = 0; w < FFR_NWARPS; ++w)
                {
                    const unsigned cnt = s_wcnt[w][lane];
                    if (w < warp) before += cnt;  /* same-key slots in earlier warps */
                    total += cnt;
                }
                unsigned incl = total;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1)
                {
                    const unsigned v = __shfl_up_sync(0xffffffffu,incl,o);
                    if (lane >= o) incl += v;
                }
                const unsigned first = incl - total + before; /* first position of (key=lane, this warp) */
                if (warp == 0 && it >= 0 && (uint32_t)lane < nx)
                    xfc += total;                 /* ++xf_dist[xf_id], buffer_renderer.hpp:172 */
                const unsigned mypos = __shfl_sync(0xffffffffu,first,key) + rank;
                s_perm[mypos] = (unsigned short)tid;
            }
            __syncthreads();
            /* ---- B: advance slot perm[tid] ---- */
            const int s = s_perm[tid];
            const uint32_t xi = s_xi[s];
            if (xi < nx)
            {
                const DevXFormT<T> &xf = xfs[xi];
                T p[D], pf[D];
                T c[RCAP > 0 ? RCAP : 1], cf[RCAP > 0 ? RCAP : 1];
#pragma unroll
                for (int i = 0; i < D; ++i)
                    p[i] = sp[i*FFR_TPB + s];
                RngT<T> wr;
                wr.col = rng_base + s;
                wr.rcol = rsl_base + s;
                const bool xr = (xf.flags & XF_USES_RNG) || (it >= 0 && has_final && (xfs[nx].flags & XF_USES_RNG));
                if (xr)
                    LOAD_RNG(wr,s);
                /* RenderIterator::iterate, render_iterator.hpp:106-139 */
                xform_apply<T,D,false>(xf,vars,wr,p,p);
                if (it >= 0)
                {
                    if (RCAP > 0)
                    {
                        FOR_COLOR(i)
                            c[i] = sc[i*FFR_TPB + s];
                        if (xf.flags & XF_HAS_COLOR)
                        {
                            const T cs = xf.color_speed;
                            FOR_COLOR(i)
                                c[i] = (1.0-cs)*c[i] + cs*__ldg(colors + xf.color_off + i);
                        }
                    }
                    if (has_final)
                    {
                        const DevXFormT<T> &xff = xfs[nx];
                        xform_apply<T,D,false>(xff,vars,wr,p,pf);
                        if (RCAP > 0)
                        {
                            if (xff.flags & XF_HAS_COLOR)
                            {
                                const T cs = xff.color_speed;
                                FOR_COLOR(i)
                                    cf[i] = (1.0-cs)*c[i] + cs*__ldg(colors + xff.color_off + i);
                            }
                            else
                            {
                                FOR_COLOR(i)
                                    cf[i] = c[i];
                            }
                        }
                    }
                    else
                    {
#pragma unroll
                        for (int i = 0; i < D; ++i)
                            pf[i] = p[i];
                        if (RCAP > 0)
                        {
                            FOR_COLOR(i)
                                cf[i] = c[i];
                        }
                    }
                    /* _render_batch body, buffer_renderer.hpp:171-229 */
                    ++n_iter;
                    bool bad = false;
                    bool gone = false;
#pragma unroll
                    for (int i = 0; i < D; ++i)
                        bad |= bad_value(p[i]);
                    if (bad) /* :175-186 */
                    {
                        u64 idx = atomicAdd(&prm.stats->n_bad,1ULL);
                        if (idx < FFR_MAX_BAD_RECORDED)
                        {
                            prm.stats->bad_xf[idx] = xf.json_id;
#pragma unroll
                            for (int i = 0; i < D; ++i)
                                prm.stats->bad_pt[idx][i] = (double)p[i];
                        }
                        if (idx + 1 > prm.bv_limit)
                        {
                            prm.stats->abort = 1;
                            gone = true;
                            s_xi[s] = 0xfffffffeu;   /* tell the owner this chain stopped */
                        }
                        else
                        {
                            /* iter.init() on the slot's own stream; pf, cf stay stale (Q3) */
                            if (!xr)
                                LOAD_RNG(wr,s);
                            ChainState<T,D,RCAP> st = chain_reinit<T,D,RCAP,false>(fl,wr);
                            wr = st.rng;
                            STORE_RNG(wr,s);
#pragma unroll
                            for (int i = 0; i < D; ++i)
                                p[i] = st.p[i];
                            FOR_COLOR(i)
                                c[i] = st.c[i];
                        }
                    }
                    else if (xr)
                        STORE_RNG(wr,s);
                    if (RCAP > 0)
                    {
                        FOR_COLOR(i)
                            sc[i*FFR_TPB + s] = c[i];
                    }
                    if (!gone)
                    {
#pragma unroll
                        for (int i = 0; i < D; ++i) /* :188-194 */
                        {
                            pmin[i] = (p[i] < pmin[i]) ? p[i] : pmin[i];
                            pmax[i] = (p[i] > pmax[i]) ? p[i] : pmax[i];
                        }
                        bool inb = true;
#pragma unroll
                        for (int i = 0; i < D; ++i)
                            inb &= (pf[i] >= fl->lo[i]) && (pf[i] <= fl->hi[i]);
                        if (inb)
                        {
                            ++n_plot;
                            u64 bi = to_index((pf[0] - fl->lo[0]) * fl->mult_d[0]);
#pragma unroll
                            for (int i = 1; i < D; ++i)
                                bi += to_index((pf[i] - fl->lo[i]) * fl->mult_d[i]) * fl->mult_i[i];
                            if (trace)
                                trace[(u64)it*prm.chain_count + (g*FFR_TPB + s)] = bi;
                            W *cell = buffer + bi*cellsz;
                            if (warp_agg)
                            {
                                const unsigned pe = __match_any_sync(__activemask(),bi);
                                if ((int)(__ffs(pe) - 1) == lane)
                                    hist_add(cell,(unsigned)__popc(pe));
                            }
                            else if (!discard)
                                hist_add(cell,1u);
                            if (RCAP > 0 && !discard)
                            {
                                FOR_COLOR(i)
                                    atomicAdd((T*)(cell + 1 + i),cf[i]);
                            }
                        }
                    }
                }
                else if (xr)
                    STORE_RNG(wr,s);
#pragma unroll
                for (int i = 0; i < D; ++i)
                    sp[i*FFR_TPB + s] = p[i];
            }
            __syncthreads();
        }
    }
#undef LOAD_RNG
#undef STORE_RNG

    /* merge statistics, buffer_renderer.hpp:232-246 */
    if (warp == 0 && (uint32_t)lane < nx && xfc)
        atomicAdd(&prm.stats->xf_dist[lane],xfc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        n_iter += __shfl_xor_sync(0xffffffffu,n_iter,o);
        n_plot += __shfl_xor_sync(0xffffffffu,n_plot,o);
#pragma unroll
        for (int i = 0; i < D; ++i)
        {
            T a = __shfl_xor_sync(0xffffffffu,pmin[i],o);
            T b = __shfl_xor_sync(0xffffffffu,pmax[i],o);
            pmin[i] = (a < pmin[i]) ? a : pmin[i];
            pmax[i] = (b > pmax[i]) ? b : pmax[i];
        }
    }
    if (lane == 0)
    {
        if (n_iter) atomicAdd(&prm.stats->s_iter,n_iter);
        if (n_plot) atomicAdd(&prm.stats->s_plot,n_plot);
#pragma unroll
        for (int i = 0; i < D; ++i)
        {
            atomicMin(&prm.stats->pt_min[i],f64_to_ordered((double)pmin[i]));
            atomicMax(&prm.stats->pt_max[i],f64_to_ordered((double)pmax[i]));
        }
    }
}

#define FFR_SMEM_REGROUP_BYTES(D,RCAP,EB) ((16 + 4 + (D) + (RCAP))*FFR_TPB*(EB))

/* K2b: fold K1e's accumulation tile into the buffer. The tile holds the counts of one launch
   with cell i at position (((i >> g)*mul) mod 2^(k-g)) << g | (i mod 2^g) -- groups of 2^g cells
   (g = 2: one 32-byte sector of u64 counts) stay together and the groups are scrambled
   multiplicatively (mul odd: a bijection)
   that spreads an attractor's strongly patterned cell addresses evenly over the L2 slices
   (measured: sierpinski@1024^2 1.40e11 -> 1.94e11 REDs/s, the uniform-random rate). The tile is
   read linearly (coalesced); most of it is zero for a sparse attractor, and only nonzero
   positions touch the buffer (the inverse map uses mul's inverse mod 2^(k-g)). The tile is left
   all zero. */
template <typename W>
__global__ void fold_acc_kernel(W *__restrict__ acc, W *__restrict__ buffer, u64 cells, uint32_t mul_inv,
        uint32_t gran)
{
    const u64 stride = (u64)gridDim.x*blockDim.x;
    const uint32_t mask = (uint32_t)((cells - 1) >> gran), low = (1u << gran) - 1u;
    for (u64 j = (u64)blockIdx.x*blockDim.x + threadIdx.x; j < cells; j += stride)
    {
        const W v = acc[j];
        if (v)
        {
            const uint32_t g = (uint32_t)(j >> gran);
            buffer[((u64)((g*mul_inv) & mask) << gran) | ((uint32_t)j & low)] += v;
            acc[j] = 0;
        }
    }
}

/* K2c: fold K1e's COMPACT tile into the buffer. For buffers whose span exceeds what the GPU's
   address translation covers (measured: a fixed 8 MB hot set takes 1.9e11 RED/s while its span
   is <= 256 MiB and 4.3-5.4e10 at 512 MiB-1 GiB, tools/micro/red_span.cu) the kernel scatters
   into rows of 512 cells allocated on first touch in a tile of <= 192 MiB; dir[row] is the
   row's slot. One warp per row; the tile is left all zero, the directory stays (rows keep their
   slots for the context's lifetime). */
/* geometry of BLOCK rows: blk[d] = log2 of a row's extent along axis d (sum FFR_DIR_ROW_SHIFT),
   rows[d] = blocks along axis d, mult[d] = the reference's index multiplier of axis d. All blk
   zero: a row is 512 consecutive cells. */
struct DirGeom
{
    uint32_t blk[3], rows[3];
    u64 mult[3];
};

template <typename W>
__global__ void fold_dir_kernel(W *__restrict__ tile, W *__restrict__ buffer, const unsigned int *__restrict__ dir,
        u64 rows, const DirGeom g)
{
    const u64 warps = ((u64)gridDim.x*blockDim.x) >> 5;
    const unsigned lane = threadIdx.x & 31u;
    const bool blocked = (g.blk[0] | g.blk[1] | g.blk[2]) != 0;
    for (u64 row = ((u64)blockIdx.x*blockDim.x + threadIdx.x) >> 5; row < rows; row += warps)
    {
        const unsigned int slot = dir[row];
        if (slot >= FFR_DIR_DIRECT)
            continue;
        W *t = tile + ((u64)slot << FFR_DIR_ROW_SHIFT);
        u64 base = row << FFR_DIR_ROW_SHIFT;
        if (blocked)
        {
            const u64 r0 = row % g.rows[0], r12 = row / g.rows[0];
            const u64 r1 = r12 % g.rows[1], r2 = r12 / g.rows[1];
            base = (r0 << g.blk[0])*g.mult[0] + (r1 << g.blk[1])*g.mult[1] + (r2 << g.blk[2])*g.mult[2];
        }
#pragma unroll 4
        for (unsigned i = lane; i < (1u << FFR_DIR_ROW_SHIFT); i += 32u)
        {
            const W v = t[i];
            if (v)
            {
                u64 at = base + i;
                if (blocked)
                {
                    const unsigned o0 = i & ((1u << g.blk[0]) - 1u);
                    const unsigned o1 = (i >> g.blk[0]) & ((1u << g.blk[1]) - 1u);
                    const unsigned o2 = i >> (g.blk[0] + g.blk[1]);
                    at = base + o0*g.mult[0] + o1*g.mult[1] + o2*g.mult[2];
                }
                buffer[at] += v;
                t[i] = 0;
            }
        }
    }
}

/* K2: dst += src with the reference's mixed element typing: element 0 of each cell is a
   count (u64 / u32), elements 1..r are colour sums (f64 / f32) (buffer_renderer.hpp:375-391).
   src may be peer memory (multi-GPU reduce over NVLink) or a staged host buffer (-i). */
template <typename T>
__global__ void add_buffer_kernel(typename Real<T>::word *__restrict__ dst,
        const typename Real<T>::word *__restrict__ src, u64 n_elems, uint32_t cellsz, u64 first_elem)
{
    typedef typename Real<T>::word W;
    const u64 stride = (u64)gridDim.x*blockDim.x;
    /* four independent loads in flight per thread (src may be a peer's memory over NVLink) */
    for (u64 i0 = (u64)blockIdx.x*blockDim.x + threadIdx.x; i0 < n_elems; i0 += 4*stride)
    {
        W s[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            const u64 i = i0 + (u64)k*stride;
            s[k] = i < n_elems ? __ldcs(src + i) : (W)0;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            const u64 i = i0 + (u64)k*stride;
            if (i >= n_elems)
                continue;
            if (cellsz == 1 || (first_elem + i) % cellsz == 0)
                dst[i] += s[k];
            else
            {
                T a, b;
                W d = dst[i];
                memcpy(&a,&d,sizeof(T));
                memcpy(&b,&s[k],sizeof(T));
                a += b;
                memcpy(&d,&a,sizeof(T));
                dst[i] = d;
            }
        }
    }
}

/* K2d: the multi-GPU exchange step as a reduce-scatter over peer memory. The device that owns
   elements [first, first + n) of the buffer adds the same slice of every other device's private
   buffer to its own, reading the peers directly over NVLink (or, under one process per GPU, the
   slices an all-to-all delivered), typed by position like K2 (element 0 of a cell is a count, the
   rest are colour sums). All devices run this at once on different slices, so every NVLink
   port carries 1/N of the buffer per peer instead of device 0 pulling N-1 whole buffers. */
#define FFR_MAX_PEERS 15
struct PeerSlices { const void *src[FFR_MAX_PEERS]; };

template <typename T>
__global__ void __launch_bounds__(256) reduce_slices_kernel(typename Real<T>::word *__restrict__ dst,
        const PeerSlices peers, int n_src, u64 first_elem, u64 n_elems, uint32_t cellsz)
{
    typedef typename Real<T>::word W;
    const u64 stride = (u64)gridDim.x*blockDim.x;
    for (u64 i = (u64)blockIdx.x*blockDim.x + threadIdx.x; i < n_elems; i += stride)
    {
        W v[FFR_MAX_PEERS];
#pragma unroll
        for (int k = 0; k < FFR_MAX_PEERS; ++k)
            if (k < n_src)
                v[k] = __ldcs((const W*)peers.src[k] + i);     /* read once, keep out of the caches */
        W d = dst[i];
        if (cellsz == 1 || (first_elem + i) % cellsz == 0)
        {
#pragma unroll
            for (int k = 0; k < FFR_MAX_PEERS; ++k)
                if (k < n_src)
                    d += v[k];
        }
        else
        {
            T a;
            memcpy(&a,&d,sizeof(T));
#pragma unroll
            for (int k = 0; k < FFR_MAX_PEERS; ++k)
                if (k < n_src)
                {
                    T b;
                    memcpy(&b,&v[k],sizeof(T));
                    a += b;
                }
            memcpy(&d,&a,sizeof(T));
        }
        dst[i] = d;
    }
}

/* K3: histogramSum / histogramMax, buffer_renderer.hpp:483-509 */
template <typename W>
__global__ void hist_sum_max_kernel(const W *__restrict__ buf, u64 cells, uint32_t cellsz,
        u64 *out_sum, u64 *out_max)
{
    u64 sum = 0, mx = 0;
    const u64 stride = (u64)gridDim.x*blockDim.x;
    for (u64 i = (u64)blockIdx.x*blockDim.x + threadIdx.x; i < cells; i += stride)
    {
        const u64 v = buf[i*cellsz];
        sum += v;
        mx = v > mx ? v : mx;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        sum += __shfl_xor_sync(0xffffffffu,sum,o);
        const u64 m2 = __shfl_xor_sync(0xffffffffu,mx,o);
        mx = m2 > mx ? m2 : mx;
    }
    if ((threadIdx.x & 31) == 0)
    {
        atomicAdd(out_sum,sum);
        atomicMax(out_max,mx);
    }
}

/* K4a: histogram min/max for the tone map (ImageRenderer::getValueBounds,
   image_renderer.hpp:112-127, with func = count) */
template <typename W>
__global__ void hist_min_max_kernel(const W *__restrict__ buf, u64 cells, uint32_t cellsz,
        u64 *out_min, u64 *out_max)
{
    u64 mn = ~0ULL, mx = 0;
    const u64 stride = (u64)gridDim.x*blockDim.x;
    for (u64 i = (u64)blockIdx.x*blockDim.x + threadIdx.x; i < cells; i += stride)
    {
        const u64 v = buf[i*cellsz];
        mn = v < mn ? v : mn;
        mx = v > mx ? v : mx;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        const u64 a = __shfl_xor_sync(0xffffffffu,mn,o);
        const u64 b = __shfl_xor_sync(0xffffffffu,mx,o);
        mn = a < mn ? a : mn;
        mx = b > mx ? b : mx;
    }
    if ((threadIdx.x & 31) == 0)
    {
        atomicMin(out_min,mn);
        atomicMax(out_max,mx);
    }
}

/* K4b: log-density tone map, one thread per cell, coalesced: reads (1+r)*sizeof(T) B, writes
   channels*bits/8 B per cell (HBM bound). Pixel math of render_image, src/ffr_img.cpp, in
   num_t arithmetic (std::log / std::pow pick the float overloads in the float build):
     gray  l = log(1+n)/max; v = pow(l,1/gamma); pix = (pix_t)(v*pix_scale)      :236-243
     rgb   v * colour_i/n per channel                                            :283-294
     mono  n != 0 ? 1 : 0                                                        :259-263
   pix_scale = 2^bits * (1 - machine eps) (constants.hpp:59-91). `max` = log(1 + hist_max)
   evaluated with the SAME device log as the cells so the brightest cell maps to exactly 1.0 as
   in the reference (log is monotonic, so this equals the max over cells of log(1+n)). Cells
   with n == 0 in RGB mode are NaN in the reference (0/0, then an undefined cast that yields 0
   on x86-64): they are 0 here. Values are clamped to the top code instead of wrapping. */
template <typename T, typename PIX>
__global__ void tonemap_kernel(const typename Real<T>::word *__restrict__ buf, u64 cells,
        uint32_t cellsz, int mode, u64 hist_max, T gp, PIX *__restrict__ out)
{
    typedef typename Real<T>::word W;
    const T adjust = sizeof(T) == 8 ? (T)(1.0 - 1.0/4503599627370496.0) : (T)(1.0f - 1.0f/8388608.0f);
    const T pix_scale = (T)(1ULL << (8*sizeof(PIX))) * adjust;
    const T top = (T)((1ULL << (8*sizeof(PIX))) - 1);
    const T mx = log(1 + (T)hist_max);
    const u64 stride = (u64)gridDim.x*blockDim.x;
    for (u64 i = (u64)blockIdx.x*blockDim.x + threadIdx.x; i < cells; i += stride)
    {
        const W *cell = buf + i*cellsz;
        const W n = cell[0];
        if (mode == FFR_TONE_MONO)
        {
            out[i] = (PIX)((n != 0 ? (T)1.0 : (T)0.0) * pix_scale);
            continue;
        }
        const T l = log(1 + (T)n) / mx;
        const T ll = pow(l,gp);
        if (mode == FFR_TONE_GRAY)
        {
            const T v = ll * pix_scale;
            out[i] = (PIX)(v > top ? top : v);
        }
        else
        {
            const T h = (T)n;
#pragma unroll
            for (int c = 0; c < 3; ++c)
            {
                T col;
                const W bits = cell[1+c];
                memcpy(&col,&bits,sizeof(T));
                T v = (n == 0) ? (T)0.0 : (ll * (col / h)) * pix_scale;
                v = v > top ? top : (v < (T)0.0 ? (T)0.0 : v);
                out[i*3+c] = (PIX)v;
            }
        }
    }
}

/* T1: one XForm::applyIteration per point with its own seeded stream */
template <typename T, int D>
__global__ void __launch_bounds__(FFR_TPB) iterate_points_kernel(const void *blob,
        uint32_t blob_bytes, int xf_slot, u64 n, const u64 *seeds, const double *pin, double *pout)
{
    typedef typename Real<T>::word W;
    extern __shared__ __align__(16) unsigned char smem[];
    W *rng_base = (W*)smem;
    DevFlameT<T> *fl = (DevFlameT<T>*)(smem + FFR_SMEM_RNG_BYTES_W(sizeof(W)));
    stage_blob(fl,blob,blob_bytes);
    const u64 i = (u64)blockIdx.x*FFR_TPB + threadIdx.x;
    if (i >= n)
        return;
    RngT<T> rng;
    rng.bind(rng_base,threadIdx.x);
    rng.seed(seeds[i]);
    T p[D];
#pragma unroll
    for (int d = 0; d < D; ++d)
        p[d] = (T)pin[i*D+d];
    xform_apply<T,D,false>(blob_xforms(fl)[xf_slot],blob_vars(fl),rng,p,p);
#pragma unroll
    for (int d = 0; d < D; ++d)
        pout[i*D+d] = (double)p[d];
}

/* T1: raw ISAAC stream (words widened to u64) */
template <typename T>
__global__ void __launch_bounds__(FFR_TPB) isaac_words_kernel(u64 seed, u64 n, u64 *out)
{
    typedef typename Real<T>::word W;
    extern __shared__ __align__(16) unsigned char smem[];
    RngT<T> rng;
    rng.bind((W*)smem,threadIdx.x);
    rng.seed(seed + threadIdx.x);
    if (threadIdx.x == 0)
        for (u64 i = 0; i < n; ++i)
            out[i] = (u64)rng.next();
}

/* M1b: attractor replay -- the scatter of a render WITHOUT the chaos game in front of it: bare
   REDs at exactly the addresses a render of this flame scattered to (FFR_SCATTER_TRACE records
   them: the buffer cell, or -- bit 63 set -- the cell of the kernel's own accumulation tile after
   the scramble / row directory). This is the measured ceiling of that render's scatter step: the
   render can approach it but not beat it.
   Built to SATURATE the atomic units rather than to wait on its trace: 2048 resident threads per
   SM, each with 8 independent 128-bit trace loads in flight and then 16 REDs (which return
   nothing, so nothing waits on them). Entry j of iteration `it` belongs to chain j, so a warp
   replays 64 chains at the same iteration and collides on hot cells as the render's warps do. */
#define FFR_REPLAY_TPB 256
#define FFR_REPLAY_UNROLL 8
template <typename T>
__global__ void __launch_bounds__(FFR_REPLAY_TPB,8) atomic_replay_kernel(typename Real<T>::word *buffer,
        typename Real<T>::word *tile, const ulonglong2 *__restrict__ trace, u64 n_pairs, uint32_t cellsz)
{
    typedef typename Real<T>::word W;
    const u64 stride = (u64)gridDim.x*FFR_REPLAY_TPB*FFR_REPLAY_UNROLL;
    for (u64 base = (u64)blockIdx.x*FFR_REPLAY_TPB*FFR_REPLAY_UNROLL + threadIdx.x; base < n_pairs; base += stride)
    {
        ulonglong2 e[FFR_REPLAY_UNROLL];
#pragma unroll
        for (int j = 0; j < FFR_REPLAY_UNROLL; ++j)
        {
            const u64 at = base + (u64)j*FFR_REPLAY_TPB;
            e[j] = at < n_pairs ? __ldcs(&trace[at]) : make_ulonglong2(~0ULL,~0ULL);
        }
#pragma unroll
        for (int j = 0; j < FFR_REPLAY_UNROLL; ++j)
        {
#pragma unroll
            for (int h = 0; h < 2; ++h)
            {
                const u64 bi = h ? e[j].y : e[j].x;
                if (bi == ~0ULL)
                    continue;
                W *cell = (bi >> 63) ? tile + (bi & ~(1ULL << 63)) : buffer + bi*cellsz;
                hist_add(cell,1u);
                for (uint32_t i = 1; i < cellsz; ++i)
                    atomicAdd((T*)(cell + i),(T)0.5);
            }
        }
    }
}

/* M1c: the same replay from a shared-memory window. The streaming replay above pays one L2
   sector read per four trace entries, which for an L2-resident scatter is a fifth of the L2's
   sector-operation rate -- the very resource being measured. Here every block copies a window of
   the trace (`win` entries, consecutive chains at consecutive iterations) into shared memory
   once and replays it `reps` times, so the timed loop issues nothing but REDs. For a scatter
   that lives in HBM the cyclic revisits distort the L2 hit rate instead, which is why the bench
   reports the better of the two replays as the ceiling. n_red receives the REDs issued. */
template <typename T>
__global__ void __launch_bounds__(1024,2) atomic_replay_window_kernel(typename Real<T>::word *buffer,
        typename Real<T>::word *tile, const u64 *__restrict__ trace, u64 n_entries, uint32_t win, uint32_t reps,
        uint32_t cellsz, u64 *n_red)
{
    typedef typename Real<T>::word W;
    extern __shared__ __align__(16) unsigned char smem[];
    u64 *w = (u64*)smem;
    const u64 base = (u64)blockIdx.x*win;
    unsigned valid = 0;
    for (uint32_t i = threadIdx.x; i < win; i += blockDim.x)
    {
        const u64 e = base + i < n_entries ? __ldcs(&trace[base + i]) : ~0ULL;
        w[i] = e;
        valid += e != ~0ULL;
    }
    __syncthreads();
    for (uint32_t r = 0; r < reps; ++r)
    {
#pragma unroll 8
        for (uint32_t i = threadIdx.x; i < win; i += blockDim.x)
        {
            const u64 bi = w[i];
            if (bi == ~0ULL)
                continue;
            W *cell = (bi >> 63) ? tile + (bi & ~(1ULL << 63)) : buffer + bi*cellsz;
            hist_add(cell,1u);
            for (uint32_t k = 1; k < cellsz; ++k)
                atomicAdd((T*)(cell + k),(T)0.5);
        }
    }
    valid = __reduce_add_sync(0xffffffffu,valid);
    if ((threadIdx.x & 31u) == 0 && valid)
        atomicAdd(n_red,(u64)valid*reps);
}

/* M1: random-atomic microbenchmark: same RED mix as the render kernel's scatter (1 count +
   r colour sums per cell) at uniformly random cells of the buffer, no chaos game in front of
   it; same launch shape as the replay (2048 threads per SM, REDs issued back to back). A
   reference point for the memory system at this buffer size, NOT a ceiling for a render: an
   attractor concentrates its samples on few cells and can stay L2-resident in a buffer that
   uniform addresses stream from HBM. */
template <typename T>
__global__ void __launch_bounds__(FFR_REPLAY_TPB,8) atomic_bench_kernel(typename Real<T>::word *buffer, u64 cells,
        uint32_t cellsz, u64 per_thread, u64 seed)
{
    typedef typename Real<T>::word W;
    u64 s = splitmix64(seed + (u64)blockIdx.x*blockDim.x + threadIdx.x);
#pragma unroll 8
    for (u64 k = 0; k < per_thread; ++k)
    {
        /* xorshift64* */
        s ^= s >> 12;
        s ^= s << 25;
        s ^= s >> 27;
        const u64 rnd = s * 0x2545F4914F6CDD1DULL;
        const u64 bi = __umul64hi(rnd,cells);
        W *cell = buffer + bi*cellsz;
        hist_add(cell,1u);
        for (uint32_t i = 1; i < cellsz; ++i)
            atomicAdd((T*)(cell + i),(T)0.5);
    }
}
