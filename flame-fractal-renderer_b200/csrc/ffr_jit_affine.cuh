/*
ffr_jit_affine.cuh -- K1e, the flame-specialised render kernel for PURE-AFFINE flames (every
variation is `linear`, no final xform, no colour dimensions): sierpinski_triangle,
barnsley_fern, sierpinski_triangle_3d, rectangle*. sm_100a, compiled at run time with NVRTC;
the generated part comes from jit::generate_affine (ffr_jit_host.cuh).

Chain semantics are those of every other render kernel here: BufferRenderer::_render_batch
(renderers/buffer_renderer.hpp:150-250) over RenderIterator::_init / iterate
(renderers/render_iterator.hpp:52-60,106-139), chain k seeded with splitmix64(base_seed + k).

Why a kernel of its own. All xforms of such a flame run the SAME code on different
coefficients, so there is no divergence to schedule around (K1b/K1d's business); what the
interpreter kernel K1 spends its 260 instructions per iteration on is interpretation:
flag tests, loops over variation lists, coefficient loads, 64-bit index arithmetic, the
int -> fp conversion of the selection draw, unconditional min/max selects. Here:

 * one thread = one chain; a WARP takes 32 consecutive chains from the launch's global
   counter -- no block-level barrier, no lock step beyond the warp;
 * the xform is one generated straight-line expression per output coordinate in which every
   coefficient that is the same for all xforms is a constant-bank operand (or folded away, see
   the exactness rules in ffr_jit_host.cuh) and only the coefficients that differ are fetched
   from a small shared-memory table [pair][xform] with one 16-byte load per pair;
 * gen() (isaac.hpp:77-90) evaluates Flame::getRandomXForm (types/flame.hpp:212-219) for each
   of its 16 result words on the spot, as integer compares on the word, and returns the 16
   selections packed 4 bits each: the per-iteration draw is one funnel shift;
 * in-bounds implies not-a-bad-value (bounds are checked to lie within the bad value
   threshold), so the bad value test runs only for samples that are not plotted;
 * the extremes (buffer_renderer.hpp:188-194) are updated behind one combined, rarely taken
   branch instead of 4*D unconditional compare-and-selects;
 * s_iter follows from the chain lengths; s_plot = s_iter - (samples not plotted);
 * for L2-resident power-of-two buffers the launch scatters into a tile of its own whose cell
   order is scrambled (JACC_MUL), folded into the buffer afterwards (fold_acc_kernel): an
   attractor's cell addresses are strongly patterned and load the L2 slices unevenly;
 * for buffers beyond the reach of the address translation (> 256 MiB) the launch scatters into
   a compact tile of 4 KiB rows allocated on first touch through a row directory (JDIR_CAP,
   fold_dir_kernel): a sparse attractor's hot rows then span tens of MiB instead of 1 GiB. A row
   is a BLOCK of 512 cells (8x8x8 in 3-d, 32x16 in 2-d; JDIR_BLOCKED) where the sizes allow it.

Everything else is the reference's arithmetic in the reference's order (-fmad=false), so the
histogram counts and statistics equal K1's bit for bit (tests/test_gpu_jit.py).

The generated translation unit defines before including this file:
  JT, JD, JNX, JTPB, JMINB, FFR_TPB (== JTPB), JNPAIR, JIDX (index type),
  jc[] (constants), jtab[] (varying coefficients, [pair][xform][2]),
  jaf_xform(tb,pin,pout), jit_select(r), jit_select_word(w), jit_inb(pf), jaf_index(pf),
  jit_json_id(k)
*/

#pragma once

#include "ffr_params.cuh"

typedef Real<JT>::word JW;

struct JafGen { JW a, b; unsigned long long keys; };

struct JafKeyHook
{
    unsigned long long keys;
    __device__ __forceinline__ void operator()(int i, u64 w)
    {
        keys |= (unsigned long long)jit_select_word((JW)w) << (4*i);
    }
};

/* gen() + the 16 selections of the new block */
__device__ __noinline__ JafGen jaf_gen(JW *col, JW *rcol, JW aa, JW bb)
{
    JafKeyHook h;
    h.keys = 0;
    GenOutT<JW> o = isaac_gen_body<JW>(col,rcol,aa,bb,h);
    JafGen r;
    r.a = o.a;
    r.b = o.b;
    r.keys = h.keys;
    return r;
}

/* FlameRNG::randNum from a generator word (flame_rng.hpp:67-87) */
__device__ __forceinline__ JT jaf_word_to_num(JW w)
{
    if (sizeof(JW) == 8)
        return (JT)((double)(w >> 11) * (1.0 / 9007199254740992.0));
    return (JT)((float)(w >> 8) * (1.0f / 16777216.0f));
}

/* A bad value (buffer_renderer.hpp:175-186): record it, then either give up (limit exceeded)
   or RenderIterator::init() on the chain's own stream (render_iterator.hpp:52-60). Cold and
   out of line; the generator goes through the general RngT path (words from randrsl). */
struct JafBad
{
    JW a, b, c;
    unsigned long long keys;
    int sh;
    int dead;
    JT p[JD];
};

__device__ __noinline__ JafBad jaf_bad(const RenderParams *prm, const JPAIR *tab, JW *col, JW *rcol,
        JW a, JW b, JW c, int sh, unsigned key, Pt<JT,JD> pbad)
{
    typedef JT T;
    JafBad o;
    o.a = a; o.b = b; o.c = c; o.sh = sh; o.keys = 0; o.dead = 0;
#pragma unroll
    for (int i = 0; i < JD; ++i)
        o.p[i] = pbad.v[i];
    const u64 idx = atomicAdd(&prm->stats->n_bad,1ULL);
    if (idx < FFR_MAX_BAD_RECORDED)
    {
        prm->stats->bad_xf[idx] = jit_json_id(key);
#pragma unroll
        for (int i = 0; i < JD; ++i)
            prm->stats->bad_pt[idx][i] = (double)pbad.v[i];
    }
    if (idx + 1 > prm->bv_limit)
    {
        atomicOr((unsigned int*)&prm->stats->abort,1u);
        o.dead = 1;
        return o;
    }
    RngT<T> rng;
    rng.col = col;
    rng.rcol = rcol;
    rng.a = a; rng.b = b; rng.c = c;
    rng.cnt = sh >> 2;
    T p[JD];
#pragma unroll
    for (int i = 0; i < JD; ++i)
        p[i] = 2.0*rng.num() - 1.0;
    for (int s = 0; s < Real<T>::settle_iters; ++s)
    {
        const unsigned xi = jit_select(rng.num());
        jaf_xform(tab + xi,p,p);
    }
    unsigned long long keys = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i)
        keys |= (unsigned long long)jit_select_word(rcol[i*FFR_TPB]) << (4*i);
    o.a = rng.a; o.b = rng.b; o.c = rng.c;
    o.sh = rng.cnt << 2;
    o.keys = keys;
#pragma unroll
    for (int i = 0; i < JD; ++i)
        o.p[i] = p[i];
    return o;
}

#ifdef JDIR_CAP
/* first touches of a row: allocate its slot in the compact tile (cold) */
__device__ __noinline__ unsigned jaf_dir_slow(unsigned int *dir, unsigned int *next, unsigned row)
{
    unsigned slot = __ldcg(&dir[row]);
    if (slot == FFR_DIR_EMPTY)
    {
        const unsigned old = atomicCAS(&dir[row],FFR_DIR_EMPTY,FFR_DIR_BUSY);
        if (old == FFR_DIR_EMPTY)
        {
            const unsigned mine = atomicAdd(next,1u);
            slot = mine < JDIR_CAP ? mine : FFR_DIR_DIRECT;
            atomicExch(&dir[row],slot);
        }
        else
            slot = old;     /* BUSY: this sample goes into the buffer; or the winner's slot */
    }
    return slot;
}
#endif

/* MODES = false: plain RED scatter. MODES = true: prm.scatter_mode honoured (warp-aggregated,
   discard, trace), used by the scatter diagnostics and the attractor-replay roofline. */
template <bool MODES>
__device__ __forceinline__ void jaf_render(const RenderParams &prm)
{
    typedef JT T;
    typedef JW W;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ unsigned long long s_xf[8];
    W *rng_base = (W*)smem;                       /* randmem columns, 16 words per chain */
    W *rsl_base = rng_base + 16*JTPB;             /* randrsl columns (read by cold paths only) */
    JPAIR *tab = (JPAIR*)(rsl_base + 16*JTPB);    /* varying coefficients [pair][xform] */
#ifdef JDC_SETS
    /* shared-memory cache of the row directory, JDC_SETS sets of two entries (tag << 16 | slot,
       0xffff: empty). ncu on sierpinski_3d@512^3: the per-lane directory loads were half of an
       L1TEX that ran at 90 % (one tag lookup per lane, 44 % hits); a hit here costs one 8-byte
       shared-memory load. Entries are immutable facts (a row keeps its slot for the context's
       lifetime), so the cache needs no invalidation; a row and its (set, tag) determine each
       other: set = (row ^ tag*K) mod JDC_SETS, tag = row >> JDC_SETBITS. */
    uint2 *dcache = (uint2*)(tab + JNPAIR*JNX);
    for (int i = threadIdx.x; i < (int)JDC_SETS; i += JTPB)
        dcache[i] = make_uint2(0xffffffffu,0xffffffffu);
#endif

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    for (int i = tid; i < JNPAIR*JNX; i += JTPB)
        tab[i] = ((const JPAIR*)jtab)[i];
    if (tid < 8)
        s_xf[tid] = 0;
    __syncthreads();

    W *col = rng_base + tid, *rcol = rsl_base + tid;
    W *__restrict__ buffer = (W*)prm.buffer;
#if defined(JACC_MUL) || defined(JDIR_CAP)
    W *__restrict__ acc = (W*)prm.acc;
#endif
    const bool warp_agg = MODES && prm.scatter_mode == FFR_SCATTER_WARP_AGG;
    const bool discard = MODES && (prm.scatter_mode == FFR_SCATTER_DISCARD || prm.scatter_mode == FFR_SCATTER_TRACE);
    u64 *__restrict__ trace = (MODES && prm.scatter_mode == FFR_SCATTER_TRACE) ? prm.trace : nullptr;
    const unsigned int chain_count = (unsigned int)prm.chain_count;   /* host: <= 2^30 per launch */
    const int chain_len = (int)prm.chain_len;                          /* host: < 2^31 */
    const int last_len = prm.last_len ? (int)prm.last_len : chain_len;

    /* per-thread statistics (buffer_renderer.hpp:156-160), merged at kernel end (:232-246) */
    u64 n_iter = 0, n_out = 0;
    T pmin[JD], pmax[JD];
#pragma unroll
    for (int i = 0; i < JD; ++i)
    {
        pmin[i] = INFINITY;
        pmax[i] = -INFINITY;
    }

    /* the next selection of the chain: Flame::getRandomXForm on the next generator word */
#define JAF_DRAW(key) \
        if (sh == 0) \
        { \
            ++c; \
            const JafGen g_ = jaf_gen(col,rcol,a,(W)(b + c)); \
            a = g_.a; \
            b = g_.b; \
            keys = g_.keys; \
            sh = 64; \
        } \
        sh -= 4; \
        const unsigned key = (unsigned)(keys >> sh) & 15u

    for (;;)
    {
        /* a warp takes 32 consecutive chains, unless the bad value limit was hit
           (buffer_renderer.hpp:152-153) */
        unsigned base = 0;
        if (lane == 0)
            base = (*(volatile uint32_t*)&prm.stats->abort) ? 0xffffffffu : atomicAdd(prm.work_counter,32u);
        base = __shfl_sync(0xffffffffu,base,0);
        if (base >= chain_count)
            break;
        const unsigned kk = base + lane;
        if (kk < chain_count)
        {
            const int len = (kk + 1u == chain_count) ? last_len : chain_len;
            /* rng::setSeed((u64)seed_k): isaac.hpp:267-282, init(false) :93-131, one gen() */
            W a, b, c;
            {
                RngT<T> g;
                g.col = col;
                g.rcol = rcol;
                g.seed_state(splitmix64(prm.base_seed + prm.chain_first + kk));
                a = g.a; b = g.b; c = g.c;
            }
            unsigned long long keys;
            int sh;
            {
                ++c;
                const JafGen g = jaf_gen(col,rcol,a,(W)(b + c));
                a = g.a;
                b = g.b;
                keys = g.keys;
            }
            /* RenderIterator::_init: p = randPoint (flame_rng.hpp:151-158), words 15, 14, .. */
            T p[JD];
#pragma unroll
            for (int i = 0; i < JD; ++i)
                p[i] = 2.0*jaf_word_to_num(rcol[(15 - i)*FFR_TPB]) - 1.0;
            sh = 4*(16 - JD);
            /* settle iterations (render_iterator.hpp:55-57): no statistics, no bad value test */
            for (int s = 0; s < Real<T>::settle_iters; ++s)
            {
                JAF_DRAW(key);
                jaf_xform(tab + key,p,p);
            }
            /* ++xf_dist[xf_id] (:172) as 16-bit fields of two registers, flushed to shared
               counters before a field can overflow */
            int it = 0;
            bool dead = false;
            while (it < len && !dead)
            {
                const int seg_end = (len - it > 32768) ? it + 32768 : len;
                u64 pk0 = 0, pk1 = 0;
                for (; it < seg_end; ++it)
                {
                    /* RenderIterator::iterate, render_iterator.hpp:106-139 */
                    JAF_DRAW(key);
                    jaf_xform(tab + key,p,p);
                    /* _render_batch body, buffer_renderer.hpp:171-229 */
                    if (JNX <= 4)
                        pk0 += 1ULL << (key*16u);
                    else
                    {
                        const u64 inc = 1ULL << ((key & 3u)*16u);
                        pk0 += (key < 4u) ? inc : 0ULL;
                        pk1 += (key < 4u) ? 0ULL : inc;
                    }
                    /* inclusive bounds (render_iterator.hpp:72-79), NaN is out (Q4); pf == p.
                       A point inside the bounds is not a bad value (host: |bounds| <= threshold) */
                    const bool in = jit_inb(p);
                    if (!in)
                    {
                        ++n_out;
                        bool bad = false;
#pragma unroll
                        for (int i = 0; i < JD; ++i)
                            bad |= bad_value(p[i]);
                        if (bad) /* :175-186; the stale pf (== the bad p) is out of bounds: not plotted (Q3) */
                        {
                            Pt<T,JD> pb;
#pragma unroll
                            for (int i = 0; i < JD; ++i)
                                pb.v[i] = p[i];
                            const JafBad r = jaf_bad(&prm,tab,col,rcol,a,b,c,sh,key,pb);
                            if (r.dead)
                            {
                                dead = true;
                                ++it;
                                break;
                            }
                            a = r.a; b = r.b; c = r.c;
                            keys = r.keys;
                            sh = r.sh;
#pragma unroll
                            for (int i = 0; i < JD; ++i)
                                p[i] = r.p[i];
                        }
                    }
                    /* extremes of p (:188-194): after the first samples a new extreme is rare */
                    bool ext = false;
#pragma unroll
                    for (int i = 0; i < JD; ++i)
                        ext |= (p[i] < pmin[i]) | (p[i] > pmax[i]);
                    if (ext)
                    {
#pragma unroll
                        for (int i = 0; i < JD; ++i)
                        {
                            pmin[i] = (p[i] < pmin[i]) ? p[i] : pmin[i];
                            pmax[i] = (p[i] > pmax[i]) ? p[i] : pmax[i];
                        }
                    }
                    if (in)
                    {
#ifdef JDIR_BLOCKED
                        unsigned brow, boff;
                        JIDX bi;
                        jaf_dir_split(p,brow,boff,bi); /* :202-209 + the cell's block row */
#else
                        const JIDX bi = jaf_index(p); /* :202-209 */
#endif
                        W *cell = buffer + bi;
#if defined(JACC_MUL) || defined(JDIR_CAP)
                        /* the diagnostics scatter into the buffer itself; the trace records the
                           address the render's scatter uses */
                        const bool tiled = !MODES || trace != nullptr;
#endif
#ifdef JACC_MUL
                        /* the launch's own tile, cells in scrambled order (fold_acc_kernel) */
                        if (tiled)
                            cell = acc + ((((((unsigned)bi >> JACC_GRAN)*JACC_MUL) & JACC_MASK) << JACC_GRAN) |
                                          ((unsigned)bi & ((1u << JACC_GRAN) - 1u)));
#endif
#ifdef JDIR_CAP
                        /* compact tile: rows of 512 cells allocated on first touch (fold_dir_kernel).
                           The directory is read through L1: a stale line can only say "not allocated
                           yet", which sends the lane to the coherent slow path. */
                        if (tiled)
                        {
#ifdef JDIR_BLOCKED
                            /* rows are BLOCKS of cells (8x8x8, 32x16): an attractor of fractal dimension
                               d < D touches far fewer blocks than runs of 512 consecutive cells, so the
                               hot part of the directory stays in L1 */
                            const unsigned row = brow, off = boff;
#else
                            const unsigned row = (unsigned)(bi >> FFR_DIR_ROW_SHIFT);
                            const unsigned off = (unsigned)bi & ((1u << FFR_DIR_ROW_SHIFT) - 1u);
#endif
#ifdef JDC_SETS
                            const unsigned tag = row >> JDC_SETBITS;
                            const unsigned set = (row ^ (tag*0x9E5u)) & (JDC_SETS - 1u);
                            const uint2 e = dcache[set];
                            unsigned slot = FFR_DIR_EMPTY;
                            if ((e.x >> 16) == tag && (e.x & 0xffffu) != 0xffffu)
                                slot = e.x & 0xffffu;
                            else if ((e.y >> 16) == tag && (e.y & 0xffffu) != 0xffffu)
                                slot = e.y & 0xffffu;
                            else
                            {
                                slot = __ldca(&prm.dir[row]);
                                if (slot >= FFR_DIR_DIRECT)
                                    slot = jaf_dir_slow(prm.dir,prm.dir_next,row);
                                if (slot < FFR_DIR_DIRECT)      /* remember it; the way alternates with the iteration */
                                    ((unsigned*)dcache)[2u*set + ((unsigned)it & 1u)] = (tag << 16) | slot;
                            }
#else
                            unsigned slot = __ldca(&prm.dir[row]);
                            if (slot >= FFR_DIR_DIRECT)
                                slot = jaf_dir_slow(prm.dir,prm.dir_next,row);
#endif
                            if (slot < FFR_DIR_DIRECT)
                                cell = acc + (((u64)slot << FFR_DIR_ROW_SHIFT) | off);
                        }
#endif
                        if (!MODES)
                            hist_add(cell,1u); /* :211-215 */
                        else
                        {
                            if (trace)
                            {
#if defined(JACC_MUL) || defined(JDIR_CAP)
                                const bool in_acc = cell >= acc && cell < acc + JACC_ELEMS;
                                trace[(u64)it*prm.chain_count + kk] = in_acc ? ((1ULL << 63) | (u64)(cell - acc))
                                                                             : (u64)(cell - buffer);
#else
                                trace[(u64)it*prm.chain_count + kk] = (u64)bi;
#endif
                            }
                            if (warp_agg)
                            {
                                const unsigned pe = __match_any_sync(__activemask(),(u64)bi);
                                if ((int)(__ffs(pe) - 1) == lane)
                                    hist_add(cell,(unsigned)__popc(pe));
                            }
                            else if (!discard)
                                hist_add(cell,1u);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < JNX; ++j)
                {
                    const unsigned f = (unsigned)((j < 4 ? pk0 : pk1) >> ((j & 3)*16)) & 0xffffu;
                    if (f)
                        atomicAdd(&s_xf[j],(unsigned long long)f);
                }
            }
            n_iter += (u64)it;   /* ++s_iter (:171) once per started iteration */
        }
        __syncwarp();
    }
#undef JAF_DRAW

    /* merge statistics, buffer_renderer.hpp:232-246 */
    u64 n_plot = n_iter - n_out;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        n_iter += __shfl_xor_sync(0xffffffffu,n_iter,o);
        n_plot += __shfl_xor_sync(0xffffffffu,n_plot,o);
#pragma unroll
        for (int i = 0; i < JD; ++i)
        {
            T x = __shfl_xor_sync(0xffffffffu,pmin[i],o);
            T y = __shfl_xor_sync(0xffffffffu,pmax[i],o);
            pmin[i] = (x < pmin[i]) ? x : pmin[i];
            pmax[i] = (y > pmax[i]) ? y : pmax[i];
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
    __syncthreads();
    if (tid < JNX && s_xf[tid])
        atomicAdd(&prm.stats->xf_dist[tid],s_xf[tid]);
}

/* one entry point per compilation (the diagnostics variant is compiled on demand) */
#ifndef JIT_MODES_ENTRY
extern "C" __global__ void __launch_bounds__(JTPB,JMINB) ffr_jit_render(const RenderParams prm)
{
    jaf_render<false>(prm);
}
#else
extern "C" __global__ void __launch_bounds__(JTPB,JMINB) ffr_jit_render_modes(const RenderParams prm)
{
    jaf_render<true>(prm);
}
#endif
