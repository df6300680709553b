/*
ffr_jit_kernel.cuh -- K1c, the flame-specialised render kernel (sm_100a), compiled at run time
with NVRTC from source that libffr_cuda generates for ONE flame (ffr_jit_host.cuh).

Same chain semantics as K1/K1b (ffr_kernels.cuh): BufferRenderer::_render_batch
(renderers/buffer_renderer.hpp:150-250) over RenderIterator::_init / iterate
(renderers/render_iterator.hpp:52-60,106-139), chain k seeded with splitmix64(base_seed + k).
What is different is how the work is laid out on the SM:

 * the flame is CODE, not data: every xform is a straight-line function whose affine
   coefficients, variation weights and parameters are literals (jx_<k> in the generated part),
   built from the same calc2d_body / calc_nd_body templates as the interpreter, so the
   arithmetic and its order are identical; what disappears is the opcode switch, the variation
   loop, the per-opcode calls and the blob loads (ncu on K1b: ~20 % of issued instructions and
   of stall samples were that dispatch);
 * chains are decoupled from threads. A block owns JNS chain SLOTS (JNS > JTPB threads) whose
   state (ISAAC columns, point, colour) lives in shared memory. Every iteration the slots are
   queued by the xform they drew (one queue per xform, filled with warp-aggregated shared
   atomics) and warps take CHUNKS of up to 32 slots of ONE xform from a shared counter until
   the iteration is exhausted: no warp ever executes two xform bodies (K1b: every warp that
   straddles a boundary of the sorted order does), the warps of a block balance dynamically
   instead of waiting for the slowest at three barriers per iteration, and the thread that
   advanced a slot draws its next xform and queues it for the next iteration right away, so
   there is ONE barrier per iteration. Queues are double buffered; their fill counters are
   triple buffered (consumed / being filled / being cleared).

Which thread advances a chain changes neither its stream nor its arithmetic: results equal
K1/K1b bit for bit (tests/test_gpu_jit.py).

The generated translation unit defines, before including this file:
  JT (num_t), JD (dims), JR (colour dims), JNX (xforms), JNS (slots/block), JTPB (threads/block),
  JMINB (blocks/SM), JHAS_FINAL, JANY_RNG (some xform or the final xform draws random numbers),
  jit_xform(k,pin,pout,rng), jit_final(pin,pout,rng), jit_select(r), jit_inb(pf), jit_index(pf),
  jit_color(k,c), jit_final_color(c,cf), jit_json_id(k)
*/

#pragma once

#include "ffr_params.cuh"

#define JKEY_NONE 0xffu
#define JRC (JR > 0 ? JR : 1)
typedef Real<JT>::word JW;

#define JIT_SMEM_BYTES ((size_t)JNS*((16 + 4)*sizeof(JW) + (JD + JR)*sizeof(JT)) + (size_t)2*JNX*JNS*sizeof(unsigned short))

/* append `slot` to the queue of xform `key` (JKEY_NONE: nothing to queue). Called by all 32
   lanes of a warp at a converged point: one shared atomic per distinct key in the warp. */
__device__ __forceinline__ void jq_push(unsigned short *qn, unsigned int *filln, unsigned key,
        unsigned slot, int lane)
{
    const unsigned peers = __match_any_sync(0xffffffffu,key);
    const unsigned rank = __popc(peers & ((1u << lane) - 1u));
    const int leader = __ffs(peers) - 1;
    unsigned base = 0;
    if (key < JNX && rank == 0)
        base = atomicAdd(&filln[key],(unsigned)__popc(peers));
    base = __shfl_sync(0xffffffffu,base,leader);
    if (key < JNX)
        qn[key*JNS + base + rank] = (unsigned short)slot;
}

struct JitChain
{
    RngT<JT> rng;
    JT p[JD];
    JT c[JRC];
};

/* out-of-line copy of the xform dispatch for the cold paths (per-lane xform index) */
__device__ __noinline__ Pt<JT,JD> jit_xform_cold(unsigned k, RngT<JT> *rng, Pt<JT,JD> pin)
{
    RngT<JT> g = *rng;
    Pt<JT,JD> out;
    jit_xform(k,pin.v,out.v,g);
    *rng = g;
    return out;
}

/* RenderIterator::init() after a bad value (render_iterator.hpp:52-60 via
   buffer_renderer.hpp:185): cold, out of line, by value */
__device__ __noinline__ JitChain jit_reinit(RngT<JT> rng)
{
    JitChain st;
    Pt<JT,JD> p;
#pragma unroll
    for (int i = 0; i < JD; ++i)
        p.v[i] = 2.0*rng.num() - 1.0;
    for (int s = 0; s < Real<JT>::settle_iters; ++s)
    {
        const unsigned xi = jit_select(rng.num());
        p = jit_xform_cold(xi,&rng,p);
    }
#pragma unroll
    for (int i = 0; i < JR; ++i)
        st.c[i] = rng.num();
#pragma unroll
    for (int i = 0; i < JD; ++i)
        st.p[i] = p.v[i];
    st.rng = rng;
    return st;
}

extern "C" __global__ void __launch_bounds__(JTPB,JMINB) ffr_jit_render(const RenderParams prm)
{
    typedef JT T;
    typedef JW W;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ unsigned int s_group;
    __shared__ unsigned int s_chunk[4];
    __shared__ unsigned int s_fill[3][8];
    W *rng_base = (W*)smem;                          /* randmem columns, 16 words per slot */
    W *st_a = rng_base + 16*JNS;                     /* randa, randb, randc, randcnt per slot */
    W *st_b = st_a + JNS;
    W *st_c = st_b + JNS;
    W *st_n = st_c + JNS;
    T *sp = (T*)(st_n + JNS);                        /* p[d][slot] */
    T *sc = sp + JD*JNS;                             /* c[i][slot] */
    unsigned short *q = (unsigned short*)(sc + JR*JNS);  /* [2][JNX][JNS] slot queues */
    W *rsl_base = (W*)prm.rsl_scratch + (size_t)blockIdx.x*16*JNS;  /* randrsl: L2-resident */

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    W *__restrict__ buffer = (W*)prm.buffer;
    const bool warp_agg = prm.scatter_mode == FFR_SCATTER_WARP_AGG;
    const bool discard = prm.scatter_mode == FFR_SCATTER_DISCARD || prm.scatter_mode == FFR_SCATTER_TRACE;
    u64 *__restrict__ trace = prm.scatter_mode == FFR_SCATTER_TRACE ? prm.trace : nullptr;

    u64 n_iter = 0, n_plot = 0;
    u64 xfc = 0;             /* warp 0, lane k: selections of xform k */
    T pmin[JD], pmax[JD];
#pragma unroll
    for (int i = 0; i < JD; ++i)
    {
        pmin[i] = INFINITY;
        pmax[i] = -INFINITY;
    }

    const u64 num_groups = (prm.chain_count + JNS - 1) / JNS;
    const int chain_len = (int)prm.chain_len;   /* host guarantees chain_len < 2^31 */

#define LOAD_RNG(R,slot) do { (R).a = st_a[slot]; (R).b = st_b[slot]; (R).c = st_c[slot]; \
        (R).cnt = (int)st_n[slot]; } while (0)
#define STORE_RNG(R,slot) do { st_a[slot] = (R).a; st_b[slot] = (R).b; st_c[slot] = (R).c; \
        st_n[slot] = (W)(R).cnt; } while (0)

    for (;;)
    {
        /* block-uniform: take the next group of JNS chains unless the bad value limit was hit
           (buffer_renderer.hpp:152-153) */
        if (tid == 0)
            s_group = (*(volatile uint32_t*)&prm.stats->abort) ? 0xffffffffu
                                                               : atomicAdd(prm.work_counter,1u);
        if (tid < 24)
            (&s_fill[0][0])[tid] = 0;
        if (tid >= 32 && tid < 36)
            s_chunk[tid-32] = 0;
        __syncthreads();
        const u64 g = s_group;
        if (g >= num_groups)
            break;
        /* seed the slots (rng::setSeed), p = randPoint (flame_rng.hpp:151-158), first draw */
        for (int s0 = 0; s0 < JNS; s0 += JTPB)
        {
            const int slot = s0 + tid;
            const u64 kk = g*JNS + slot;
            unsigned key = JKEY_NONE;
            if (slot < JNS && kk < prm.chain_count)
            {
                RngT<T> rng;
                rng.col = rng_base + slot;
                rng.rcol = rsl_base + slot;
                rng.seed(splitmix64(prm.base_seed + prm.chain_first + kk));
#pragma unroll
                for (int i = 0; i < JD; ++i)
                    sp[i*JNS + slot] = 2.0*rng.num() - 1.0;
                key = jit_select(rng.num());
                STORE_RNG(rng,slot);
            }
            jq_push(q,s_fill[0],key,(unsigned)slot,lane);
        }

        int b2 = 0, b3 = 0;
        /* iterations -settle..-1 are the settle iterations of _init (no stats, no plotting) */
        for (int it = -Real<T>::settle_iters; it < chain_len; ++it)
        {
            __syncthreads();   /* queue b2 complete; queue b2^1 and counters b3-1 are free */
            const int b3n = (b3 == 2) ? 0 : b3 + 1;
            const int b3p = (b3 == 0) ? 2 : b3 - 1;
            if (warp == 0 && lane < JNX && it >= 0)
                xfc += s_fill[b3][lane];              /* ++xf_dist[xf_id], buffer_renderer.hpp:172 */
            if (tid >= 32 && tid < 40)
                s_fill[b3p][tid-32] = 0;
            if (tid == 40)
                s_chunk[b3p] = 0;
            const unsigned short *qc = q + b2*(JNX*JNS);
            unsigned short *qn = q + (b2^1)*(JNX*JNS);
            for (;;)
            {
                unsigned cidx = 0;
                if (lane == 0)
                    cidx = atomicAdd(&s_chunk[b3],1u);
                cidx = __shfl_sync(0xffffffffu,cidx,0);
                /* chunk number -> (xform k, first entry, entries) */
                unsigned k = JNX, off = cidx, cnt = 0;
#pragma unroll
                for (int j = 0; j < JNX; ++j)
                {
                    const unsigned f = s_fill[b3][j];
                    const unsigned nch = (f + 31u) >> 5;
                    if (k == JNX)
                    {
                        if (off < nch)
                        {
                            k = (unsigned)j;
                            cnt = f - off*32u;
                        }
                        else
                            off -= nch;
                    }
                }
                if (k == JNX)
                    break;
                const unsigned n = cnt < 32u ? cnt : 32u;
                unsigned newkey = JKEY_NONE, slot = 0;
                if ((unsigned)lane < n)
                {
                    slot = qc[k*JNS + off*32u + lane];
                    T p[JD], pf[JD];
                    T c[JRC], cf[JRC];
#pragma unroll
                    for (int i = 0; i < JD; ++i)
                        p[i] = sp[i*JNS + slot];
                    RngT<T> rng;
                    rng.col = rng_base + slot;
                    rng.rcol = rsl_base + slot;
                    if (JANY_RNG)
                        LOAD_RNG(rng,slot);
                    /* RenderIterator::iterate, render_iterator.hpp:106-139 */
                    jit_xform(k,p,p,rng);
                    bool gone = false;
                    if (it >= 0)
                    {
                        if (JR > 0)
                        {
#pragma unroll
                            for (int i = 0; i < JR; ++i)
                                c[i] = sc[i*JNS + slot];
                            jit_color(k,c);
                        }
                        if (JHAS_FINAL)
                        {
                            jit_final(p,pf,rng);
                            if (JR > 0)
                                jit_final_color(c,cf);
                        }
                        else
                        {
#pragma unroll
                            for (int i = 0; i < JD; ++i)
                                pf[i] = p[i];
#pragma unroll
                            for (int i = 0; i < JR; ++i)
                                cf[i] = c[i];
                        }
                        if (!JANY_RNG)
                            LOAD_RNG(rng,slot);
                        /* _render_batch body, buffer_renderer.hpp:171-229 */
                        ++n_iter;
                        bool bad = false;
#pragma unroll
                        for (int i = 0; i < JD; ++i)
                            bad |= bad_value(p[i]);
                        if (bad) /* :175-186 */
                        {
                            u64 idx = atomicAdd(&prm.stats->n_bad,1ULL);
                            if (idx < FFR_MAX_BAD_RECORDED)
                            {
                                prm.stats->bad_xf[idx] = jit_json_id(k);
#pragma unroll
                                for (int i = 0; i < JD; ++i)
                                    prm.stats->bad_pt[idx][i] = (double)p[i];
                            }
                            if (idx + 1 > prm.bv_limit)
                            {
                                prm.stats->abort = 1;
                                gone = true;
                            }
                            else
                            {
                                /* iter.init() on the slot's own stream; pf, cf stay stale (Q3) */
                                JitChain st = jit_reinit(rng);
                                rng.a = st.rng.a;
                                rng.b = st.rng.b;
                                rng.c = st.rng.c;
                                rng.cnt = st.rng.cnt;
#pragma unroll
                                for (int i = 0; i < JD; ++i)
                                    p[i] = st.p[i];
#pragma unroll
                                for (int i = 0; i < JR; ++i)
                                    c[i] = st.c[i];
                            }
                        }
                        if (!gone)
                        {
#pragma unroll
                            for (int i = 0; i < JD; ++i) /* :188-194 */
                            {
                                pmin[i] = (p[i] < pmin[i]) ? p[i] : pmin[i];
                                pmax[i] = (p[i] > pmax[i]) ? p[i] : pmax[i];
                            }
                            /* inclusive bounds on pf (render_iterator.hpp:72-79); NaN is out (Q4) */
                            if (jit_inb(pf))
                            {
                                ++n_plot;
                                const u64 bi = jit_index(pf); /* :202-209 */
                                if (trace)
                                    trace[(u64)it*prm.chain_count + (g*JNS + slot)] = bi;
                                W *cell = buffer + bi*(1 + JR);
                                if (warp_agg)
                                {
                                    const unsigned pe = __match_any_sync(__activemask(),bi);
                                    if ((int)(__ffs(pe) - 1) == lane)
                                        hist_add(cell,(unsigned)__popc(pe));
                                }
                                else if (!discard)
                                    hist_add(cell,1u); /* :211-215 */
                                if (JR > 0 && !discard)
                                {
#pragma unroll
                                    for (int i = 0; i < JR; ++i) /* :217-229 */
                                        atomicAdd((T*)(cell + 1 + i),cf[i]);
                                }
                            }
                        }
                    }
                    else if (!JANY_RNG)
                        LOAD_RNG(rng,slot);
                    /* next iteration of this chain: colours are drawn after the last settle
                       iteration (render_iterator.hpp:58-59), then the xform selection */
                    const int nit = it + 1;
                    if (JR > 0 && nit == 0)
                    {
#pragma unroll
                        for (int i = 0; i < JR; ++i)
                            c[i] = rng.num();
                    }
                    const u64 kk = g*JNS + slot;
                    const int len = (kk + 1 == prm.chain_count && prm.last_len) ? (int)prm.last_len : chain_len;
                    if (!gone && nit < len)
                        newkey = jit_select(rng.num());
                    STORE_RNG(rng,slot);
#pragma unroll
                    for (int i = 0; i < JD; ++i)
                        sp[i*JNS + slot] = p[i];
                    if (JR > 0 && nit >= 0)
                    {
#pragma unroll
                        for (int i = 0; i < JR; ++i)
                            sc[i*JNS + slot] = c[i];
                    }
                }
                jq_push(qn,s_fill[b3n],newkey,slot,lane);
            }
            b2 ^= 1;
            b3 = b3n;
        }
        __syncthreads();   /* every warp is done with this group's queues and counters */
    }
#undef LOAD_RNG
#undef STORE_RNG

    /* merge statistics, buffer_renderer.hpp:232-246 */
    if (warp == 0 && lane < JNX && xfc)
        atomicAdd(&prm.stats->xf_dist[lane],xfc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        n_iter += __shfl_xor_sync(0xffffffffu,n_iter,o);
        n_plot += __shfl_xor_sync(0xffffffffu,n_plot,o);
#pragma unroll
        for (int i = 0; i < JD; ++i)
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
        for (int i = 0; i < JD; ++i)
        {
            atomicMin(&prm.stats->pt_min[i],f64_to_ordered((double)pmin[i]));
            atomicMax(&prm.stats->pt_max[i],f64_to_ordered((double)pmax[i]));
        }
    }
}
