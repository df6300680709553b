/*
ffr_jit_async.cuh -- K1d, the barrier-free form of the flame-specialised render kernel
(sm_100a, compiled at run time with NVRTC; generated part: ffr_jit_host.cuh).

Chain semantics are those of every other render kernel here: BufferRenderer::_render_batch
(renderers/buffer_renderer.hpp:150-250) over RenderIterator::_init / iterate
(renderers/render_iterator.hpp:52-60,106-139), chain k seeded with splitmix64(base_seed + k).

K1c (ffr_jit_kernel.cuh) advances all chain slots of a block in lock step: one barrier per
iteration, at which warps wait for the slowest chunk (ncu: 18 % of stall samples), and one
partly filled chunk per xform and iteration (27.5 of 32 lanes active). Here nothing is in lock
step:

 * a block owns JNS chain slots (state in shared memory) and one RING QUEUE per xform, plus one
   for "this slot's generator needs its next block of 16 words" (ISAAC gen(), isaac.hpp:77-90,
   which would otherwise run with one or two active lanes in almost every warp once the slots'
   draw counters drift apart);
 * a warp pops 32 slots from the fullest queue, advances them by ONE iteration of that xform
   (or runs gen() for them), draws each slot's next xform and pushes the slot to that xform's
   queue; a slot whose chain ended takes the next chain index from the launch's global
   counter on the spot. No barrier, no iteration counter shared between slots;
 * with JNS >= JTPB + 32*(JNX+1) there is always a queue holding >= 32 slots (at most JTPB slots
   are in flight, the rest is spread over JNX+1 queues), so warps pop full chunks until the
   launch drains.

Queues are multi-producer/multi-consumer rings of slot numbers: producers reserve positions
with one warp-aggregated atomic add on the tail and then write the entries; consumers claim a
range with a CAS on the head and wait for each claimed entry to become valid (a producer that
reserved a position writes it a few instructions later and waits for nothing in between).
At most JNS entries are ever reserved-and-unread over all queues (one per slot), so a ring of
JCAP >= JNS entries cannot overrun. An entry carries the parity of its lap around the ring in
bit 15 (position / JCAP mod 2): a consumer of position p waits until the entry shows p's lap,
so nobody ever has to mark an entry as consumed (every reserved position is written exactly
once per lap, hence the previous lap's value always has the other parity).
Memory model: the entry is published with st.release.cta and read with ld.acquire.cta, which
orders the slot's state (plain shared-memory stores before the push, plain loads after the pop)
across the two warps; head and tail are relaxed atomics. JQ_ACQREL = 0 compiles the entry
accesses as volatile instead (what round 1 shipped: correct on this hardware, where one thread's
shared-memory stores are performed in order, but outside the PTX memory model).

Which warp advances a chain, and in which order chains interleave, changes neither a chain's
stream nor its arithmetic: histogram counts and statistics equal K1/K1b/K1c bit for bit
(tests/test_gpu_jit.py).

The generated translation unit defines, before including this file, what ffr_jit_kernel.cuh
lists, plus JCAP.
*/

#pragma once

#include "ffr_params.cuh"

#ifndef JQ_ACQREL
#define JQ_ACQREL 1
#endif
#define JKEY_NONE 0xffu
#define JRC (JR > 0 ? JR : 1)
#define JNQ (JNX + 1)            /* xform queues + the gen() queue */
#define JQ_GEN JNX
#define JMASK (JCAP - 1)          /* JCAP: ring capacity, the power of two >= JNS */
#define JLAP(pos) ((((pos) & (unsigned)JCAP) != 0u) ? 0x8000u : 0u)   /* lap parity of a position */
#define JABORT_WATCHDOG 0x100u
typedef Real<JT>::word JW;

/* queue primitives on 32-bit shared-window addresses (no generic-address arithmetic, no branch
   around the one-lane atomics: they are predicated instructions) */
__device__ __forceinline__ void jq_publish(unsigned addr, unsigned v)
{
#if JQ_ACQREL
    asm volatile("st.release.cta.shared.u16 [%0], %1;" :: "r"(addr), "h"((unsigned short)v) : "memory");
#else
    asm volatile("st.volatile.shared.u16 [%0], %1;" :: "r"(addr), "h"((unsigned short)v) : "memory");
#endif
}

__device__ __forceinline__ unsigned jq_peek(unsigned addr)
{
    unsigned short v;
#if JQ_ACQREL
    asm volatile("ld.acquire.cta.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
#else
    asm volatile("ld.volatile.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
#endif
    return v;
}

/* old = CAS(addr, cmp, val) on the lanes where p is set; other lanes get ~cmp */
__device__ __forceinline__ unsigned jq_cas_pred(unsigned addr, unsigned cmp, unsigned val, bool p)
{
    unsigned old = ~cmp;
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %4, 0;\n\t@q atom.relaxed.cta.shared.cas.b32 %0, [%1], %2, %3;\n\t}"
        : "+r"(old) : "r"(addr), "r"(cmp), "r"(val), "r"((unsigned)p) : "memory");
    return old;
}

/* old = fetch-and-add on the lanes where p is set; other lanes get 0 */
__device__ __forceinline__ unsigned jq_add_pred(unsigned addr, unsigned val, bool p)
{
    unsigned old = 0;
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\t@q atom.relaxed.cta.shared.add.u32 %0, [%1], %2;\n\t}"
        : "+r"(old) : "r"(addr), "r"(val), "r"((unsigned)p) : "memory");
    return old;
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

/* FlameRNG::randNum from a generator word (flame_rng.hpp:67-87) */
__device__ __forceinline__ JT jit_word_to_num(JW w)
{
    if (sizeof(JW) == 8)
        return (JT)((double)(w >> 11) * (1.0 / 9007199254740992.0));
    return (JT)((float)(w >> 8) * (1.0f / 16777216.0f));
}

/* The 16 words of a fresh generator block decide (mostly) the next 16 xform selections of the
   chain: Flame::getRandomXForm (types/flame.hpp:212-219) is evaluated for every word where
   gen() produces it, by a converged warp, and the 16 indices are packed 4 bits each. The per
   iteration path then needs neither the word (it lives in the L2-resident scratch; slots that
   are no longer in lock step would fetch a 32-byte sector per 8-byte word) nor the integer to
   floating point conversion and the compares. Draws other than the selection (colours, random
   variations, re-initialisation) still read the words themselves. */
struct JitKeyHook
{
    unsigned long long keys;
    __device__ __forceinline__ void operator()(int i, u64 w)
    {
        keys |= (unsigned long long)jit_select_word((JW)w) << (4*i);
    }
};

struct JitGenOut { JW a, b; unsigned long long keys; };

__device__ __noinline__ JitGenOut jit_gen_keys(JW *col, JW *rcol, JW aa, JW bb)
{
    JitKeyHook h;
    h.keys = 0;
    GenOutT<JW> o = isaac_gen_body<JW>(col,rcol,aa,bb,h);
    JitGenOut r;
    r.a = o.a;
    r.b = o.b;
    r.keys = h.keys;
    return r;
}

/* the same packing from the stored block, after a draw elsewhere ran gen() inline (cold) */
__device__ __noinline__ unsigned long long jit_keys_from_rsl(const JW *rcol)
{
    unsigned long long keys = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i)
        keys |= (unsigned long long)jit_select_word(FFR_RSL_LOAD(&rcol[i*JNS])) << (4*i);
    return keys;
}

#ifndef JCOLD_START
#define JCOLD_START 0
#endif
#ifndef JCOLD_IDLE
#define JCOLD_IDLE 0
#endif
#ifndef JCOLD_BAD
#define JCOLD_BAD 0
#endif

/* A slot takes a new chain: rng::setSeed((u64)seed_k), the first block of generator words, p =
   randPoint (flame_rng.hpp:151-158). Once per chain; JCOLD_START compiles it out of line. */
struct JitStart
{
    JW a, b, c;
    unsigned long long keys;
    JT p[JD];
    int cnt;
};

__device__ __noinline__ JitStart jit_start_chain(JW *col, JW *rcol, u64 seed)
{
    RngT<JT> g;
    g.col = col;
    g.rcol = rcol;
    g.seed_state(seed);
    ++g.c;
    const JitGenOut go = jit_gen_keys(g.col,g.rcol,g.a,(JW)(g.b + g.c));
    g.a = go.a;
    g.b = go.b;
    g.cnt = 16;
    JitStart st;
#pragma unroll
    for (int i = 0; i < JD; ++i)
        st.p[i] = 2.0*g.num() - 1.0;
    --g.cnt;        /* the first selection: word 15 - JD of the block */
    st.a = g.a;
    st.b = g.b;
    st.c = g.c;
    st.keys = go.keys;
    st.cnt = g.cnt;
    return st;
}

/* A bad value (buffer_renderer.hpp:175-186): count it, record the first FFR_MAX_BAD_RECORDED, and
   say whether the limit is now exceeded (the launch then stops handing out chains). */
__device__ __noinline__ bool jit_record_bad(DevStats *stats, u64 bv_limit, u64 json_id, Pt<JT,JD> p)
{
    const u64 idx = atomicAdd(&stats->n_bad,1ULL);
    if (idx < FFR_MAX_BAD_RECORDED)
    {
        stats->bad_xf[idx] = json_id;
#pragma unroll
        for (int i = 0; i < JD; ++i)
            stats->bad_pt[idx][i] = (double)p.v[i];
    }
    if (idx + 1 > bv_limit)
    {
        atomicOr((unsigned int*)&stats->abort,1u);
        return true;
    }
    return false;
}

/* A warp found nothing to pop (converged, all 32 lanes; hd = the queue heads as the lanes read
   them, lane < JNQ, else 0): live check, back-off, watchdog. stop: no live chain is left in the
   block, or the watchdog fired. */
struct JitIdle
{
    unsigned long long spins;
    unsigned last_sig;
    int stop;
};

__device__ __noinline__ JitIdle jit_idle(const int *live_p, unsigned hd, unsigned long long spins,
        unsigned last_sig, unsigned int *abort_p)
{
    JitIdle r;
    r.spins = spins;
    r.last_sig = last_sig;
    r.stop = 0;
    int live = 0;
    if ((threadIdx.x & 31) == 0)
        live = *(volatile const int*)live_p;
    live = __shfl_sync(0xffffffffu,live,0);
    if (live == 0)
    {
        r.stop = 1;
        return r;
    }
    __nanosleep(64);
    const unsigned sig = __reduce_add_sync(0xffffffffu,hd);
    if (sig != r.last_sig)
    {
        r.last_sig = sig;
        r.spins = 0;
    }
    if (++r.spins > (1ULL << 24))
    {
        if ((threadIdx.x & 31) == 0)
            atomicOr(abort_p,JABORT_WATCHDOG);
        r.stop = 1;
    }
    return r;
}

/* MODES = false: plain RED scatter. MODES = true: prm.scatter_mode honoured (warp-aggregated,
   discard, trace), used by the scatter diagnostics and the attractor-replay roofline -- a second
   entry point, so that the render kernel proper carries neither the tests nor the code. */
template <bool MODES>
__device__ __forceinline__ void jit_render_async(const RenderParams &prm)
{
    typedef JT T;
    typedef JW W;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ __align__(8) uint2 s_ht[JNQ];         /* per queue: .x = head (claimed), .y = tail (reserved) */
    __shared__ int s_live;
    __shared__ unsigned long long s_wxf[(JTPB/32)*JNX];   /* per warp: iterations executed per xform (filled at the end) */
    W *rng_base = (W*)smem;                          /* randmem columns, 16 words per slot */
    W *st_a = rng_base + 16*JNS;                     /* randa, randb, randc, randcnt per slot */
    W *st_b = st_a + JNS;
    W *st_c = st_b + JNS;
    W *st_n = st_c + JNS;
    T *sp = (T*)(st_n + JNS);                        /* p[d][slot] */
    T *sc = sp + JD*JNS;                             /* c[i][slot] */
    unsigned long long *s_keys = (unsigned long long*)(sc + JR*JNS);  /* 16 packed selections */
    int *s_it = (int*)(s_keys + JNS);                /* iteration number of the slot's chain */
    unsigned int *s_chain = (unsigned int*)(s_it + JNS);  /* chain index within this launch */
    unsigned short *ring = (unsigned short*)(s_chain + JNS);  /* [JNQ][JCAP] */
#if JRSL_SMEM
    W *rsl_base = (W*)(ring + JNQ*JCAP);             /* randrsl columns (flames with drawing variations) */
#else
    W *rsl_base = (W*)prm.rsl_scratch + (size_t)blockIdx.x*16*JNS;  /* randrsl: L2-resident scratch */
#endif

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    W *__restrict__ buffer = (W*)prm.buffer;
    const bool warp_agg = MODES && prm.scatter_mode == FFR_SCATTER_WARP_AGG;
    const bool discard = MODES && (prm.scatter_mode == FFR_SCATTER_DISCARD || prm.scatter_mode == FFR_SCATTER_TRACE);
    u64 *__restrict__ trace = (MODES && prm.scatter_mode == FFR_SCATTER_TRACE) ? prm.trace : nullptr;
    const unsigned int chain_count = (unsigned int)prm.chain_count;   /* host: < 2^31 per launch */
    const int chain_len = (int)prm.chain_len;                          /* host: < 2^31 */
    const int last_len = prm.last_len ? (int)prm.last_len : chain_len;

    /* per-thread statistics (buffer_renderer.hpp:156-160), merged at kernel end (:232-246).
       ++s_iter and ++xf_dist[id] (:171-172) are counted once per popped chunk, below; s_iter is
       their sum. */
    u64 n_plot = 0;
    u64 wcnt = 0;      /* lane q of a warp: iterations the warp executed of xform q (JNX <= 8 < 32) */
    T pmin[JD], pmax[JD];
#pragma unroll
    for (int i = 0; i < JD; ++i)
    {
        pmin[i] = INFINITY;
        pmax[i] = -INFINITY;
    }

    for (int i = tid; i < JNQ*JCAP; i += JTPB)
        ring[i] = (unsigned short)0x8000u;     /* the lap before the first: valid for nobody */
    if (tid < JNQ)
    {
        s_ht[tid] = make_uint2(0u,0u);
    }
    const unsigned ht_s = (unsigned)__cvta_generic_to_shared(&s_ht[0]);
    const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring);
    if (tid == 0)
        s_live = JNS;
    __syncthreads();

#define LOAD_ABC(R,slot) do { (R).a = st_a[slot]; (R).b = st_b[slot]; (R).c = st_c[slot]; } while (0)
#define STORE_ABC(R,slot) do { st_a[slot] = (R).a; st_b[slot] = (R).b; st_c[slot] = (R).c; } while (0)

#if JCOLD_START
#define JIT_CHAIN_START(slot,key,kk_) do { \
        const JitStart st_ = jit_start_chain(rng_base + (slot),rsl_base + (slot), \
            splitmix64(prm.base_seed + prm.chain_first + kk_)); \
        _Pragma("unroll") \
        for (int i_ = 0; i_ < JD; ++i_) \
            sp[i_*JNS + (slot)] = st_.p[i_]; \
        (key) = (unsigned)(st_.keys >> (4*st_.cnt)) & 15u; \
        s_keys[slot] = st_.keys; \
        st_a[slot] = st_.a; \
        st_b[slot] = st_.b; \
        st_c[slot] = st_.c; \
        st_n[slot] = (W)st_.cnt; \
    } while (0)
#else
#define JIT_CHAIN_START(slot,key,kk_) do { \
        /* rng::setSeed((u64)seed_k), p = randPoint (flame_rng.hpp:151-158), first draw */ \
        RngT<T> g_; \
        g_.col = rng_base + (slot); \
        g_.rcol = rsl_base + (slot); \
        g_.seed_state(splitmix64(prm.base_seed + prm.chain_first + kk_)); \
        ++g_.c; \
        const JitGenOut go_ = jit_gen_keys(g_.col,g_.rcol,g_.a,(W)(g_.b + g_.c)); \
        g_.a = go_.a; \
        g_.b = go_.b; \
        g_.cnt = 16; \
        _Pragma("unroll") \
        for (int i_ = 0; i_ < JD; ++i_) \
            sp[i_*JNS + (slot)] = 2.0*g_.num() - 1.0; \
        --g_.cnt;       /* the first selection: word 15 - JD of the block */ \
        (key) = (unsigned)(go_.keys >> (4*g_.cnt)) & 15u; \
        s_keys[slot] = go_.keys; \
        STORE_ABC(g_,slot); \
        st_n[slot] = (W)g_.cnt; \
    } while (0)
#endif

    /* Shared tail of every step: lanes flagged `fresh` take a new chain (or retire their slot),
       then every lane with a key queues its slot. Called by all 32 lanes, converged. */
#define JIT_START_AND_PUSH(fresh,slot,key) do { \
        const unsigned fm_ = __ballot_sync(0xffffffffu,(fresh)); \
        if (fm_) \
        { \
            unsigned base_ = 0; \
            if (lane == __ffs(fm_) - 1) \
                base_ = (*(volatile uint32_t*)&prm.stats->abort) ? 0xffffffffu \
                      : atomicAdd(prm.work_counter,(unsigned)__popc(fm_)); \
            base_ = __shfl_sync(0xffffffffu,base_,__ffs(fm_) - 1); \
            if (fresh) \
            { \
                const unsigned kk_ = base_ == 0xffffffffu ? 0xffffffffu \
                                   : base_ + __popc(fm_ & ((1u << lane) - 1u)); \
                if (kk_ < chain_count) \
                { \
                    JIT_CHAIN_START(slot,key,kk_); \
                    s_it[slot] = -Real<T>::settle_iters; \
                    s_chain[slot] = kk_; \
                } \
                else \
                { \
                    (key) = JKEY_NONE; \
                    atomicSub(&s_live,1); \
                } \
            } \
        } \
        { \
            const unsigned peers_ = __match_any_sync(0xffffffffu,(key)); \
            const unsigned rank_ = __popc(peers_ & ((1u << lane) - 1u)); \
            const int leader_ = __ffs(peers_) - 1; \
            const bool queued_ = (key) < JNQ; \
            unsigned pos_ = jq_add_pred(ht_s + 8u*(key) + 4u,(unsigned)__popc(peers_),queued_ && rank_ == 0); \
            pos_ = __shfl_sync(0xffffffffu,pos_,leader_) + rank_; \
            /* The slot's state (plain shared-memory stores above) must be visible before its \
               queue entry: the entry is a release store (JQ_ACQREL), which costs one \
               MEMBAR.ALL.CTA and leaves the scatter a RED -- a fence.acq_rel or a wider membar \
               anywhere in this kernel makes ptxas turn every scatter RED into an ATOMG whose \
               (discarded) return value the warp then waits for, measured 2x on the colour \
               flames. Generator words that other warps may read go to shared memory when the \
               flame has drawing variations (JRSL_SMEM); otherwise the global scratch is only \
               read by cold paths many steps after the gen() that wrote it, with L1-bypassing \
               loads. */ \
            if (queued_) \
                jq_publish(ring_s + 2u*((key)*JCAP + (pos_ & JMASK)),(unsigned)(slot) | JLAP(pos_)); \
        } \
    } while (0)

    /* every slot starts a chain */
    for (int s0 = 0; s0 < JNS; s0 += JTPB)
    {
        const int slot = s0 + tid;
        const bool fresh = slot < JNS;
        unsigned key = JKEY_NONE;
        JIT_START_AND_PUSH(fresh,slot,key);
    }

    unsigned long long spins = 0;   /* idle polls since this BLOCK last made progress */
    unsigned last_sig = 0;
#ifdef JROT_STATIC
    /* warps that share a scheduler (warp index mod 4) prefer the same queues, so that its
       instruction cache holds fewer xform bodies */
    unsigned rot = (((unsigned)(tid >> 5) & 3u)*(unsigned)JNQ) >> 2;
#else
    unsigned rot = (unsigned)(tid >> 5);
#endif
    for (;;)
    {
        /* ---- pop up to 32 slots: the first queue holding >= 32, scanning from a position that
           rotates per warp and step so that warps spread over the queues ---- */
        unsigned q = JKEY_NONE, h0 = 0, n = 0;
        int fails = 0;
#ifndef JROT_STATIC
        ++rot;
#endif
        for (;;)
        {
            unsigned av = 0, hd = 0;
            if (lane < JNQ)
            {
                /* one 64-bit load; a consistent pair is not needed, the CAS validates head */
                const unsigned long long ht = *(volatile unsigned long long*)&s_ht[lane];
                hd = (unsigned)ht;
                av = (unsigned)(ht >> 32) - hd;
                if (av > (unsigned)JNS)
                    av = 0;      /* torn read across a concurrent pop: rescan */
            }
            unsigned m = __ballot_sync(0xffffffffu,av >= 32u);
            if (!m && fails >= 3)
                m = __ballot_sync(0xffffffffu,av > 0u);   /* drain: partly filled chunks */
            if (m)
            {
                const unsigned r = rot % (unsigned)JNQ;
                const unsigned hi = m >> r;
                const unsigned bl = hi ? (unsigned)__ffs(hi) - 1u + r : (unsigned)__ffs(m) - 1u;
                const unsigned hexp = __shfl_sync(0xffffffffu,hd,(int)bl);
                unsigned want = __shfl_sync(0xffffffffu,av,(int)bl);
                want = want < 32u ? want : 32u;
                unsigned ok = jq_cas_pred(ht_s + 8u*bl,hexp,hexp + want,lane == 0) == hexp ? 1u : 0u;
                ok = __shfl_sync(0xffffffffu,ok,0);
                if (ok)
                {
                    q = bl;
                    h0 = hexp;
                    n = want;
                    break;
                }
                continue;
            }
#if JCOLD_IDLE
            {
                const JitIdle idle = jit_idle(&s_live,hd,spins,last_sig,(unsigned int*)&prm.stats->abort);
                if (idle.stop)
                    break;
                spins = idle.spins;
                last_sig = idle.last_sig;
                ++fails;
                continue;
            }
#endif
            int live = 0;
            if (lane == 0)
                live = *(volatile int*)&s_live;
            live = __shfl_sync(0xffffffffu,live,0);
            if (live == 0)
                break;
            ++fails;
            __nanosleep(64);
            /* Watchdog against a lost slot (a protocol bug), not against waiting: with fewer live
               chains than threads -- a handful of very long chains, or the drain of a launch --
               warps legitimately find nothing to pop for as long as the launch runs. The queue
               heads move whenever ANY warp of the block pops, so only 2^24 consecutive polls
               during which no head moved count as a stall. */
            {
                const unsigned sig = __reduce_add_sync(0xffffffffu,hd);
                if (sig != last_sig)
                {
                    last_sig = sig;
                    spins = 0;
                }
            }
            if (++spins > (1ULL << 24))
            {
                if (lane == 0)
                    atomicOr((unsigned int*)&prm.stats->abort,JABORT_WATCHDOG);
                break;
            }
        }
        if (q == JKEY_NONE)
            break;
        spins = 0;      /* the watchdog measures one stall, not the idle time of a whole launch */

        const bool act = (unsigned)lane < n;
        unsigned slot = 0;
        if (act)
        {
            const unsigned pos = h0 + (unsigned)lane;
            const unsigned ea = ring_s + 2u*(q*JCAP + (pos & JMASK));
            const unsigned lap = JLAP(pos);
            unsigned v = jq_peek(ea);
            unsigned long long w = 0;
            while ((v & 0x8000u) != lap && ++w < (1ULL << 28))
                v = jq_peek(ea);
            if ((v & 0x8000u) != lap)
            {
                atomicOr((unsigned int*)&prm.stats->abort,JABORT_WATCHDOG);
                v = 0;
            }
            slot = v & 0x7fffu;
        }

        unsigned newkey = JKEY_NONE;
        bool fresh = false;
        int it = -1;
        if (q != JQ_GEN)
        {
            /* ++s_iter, ++xf_dist[xf_id] (buffer_renderer.hpp:171-172): the whole chunk runs xform q,
               so one popc per chunk instead of JNX compare-adds per lane and iteration */
            if (act)
                it = s_it[slot];
            const unsigned cm = __ballot_sync(0xffffffffu,it >= 0);
            if ((unsigned)lane == q)
                wcnt += (unsigned long long)__popc(cm);   /* lane q counts xform q: one predicated add */
        }
        if (q == JQ_GEN)
        {
            /* the slot's select draw found randcnt == 0: next(), isaac.hpp:321-329 */
            if (act)
            {
                RngT<T> rng;
                rng.col = rng_base + slot;
                rng.rcol = rsl_base + slot;
                LOAD_ABC(rng,slot);
                ++rng.c;
                const JitGenOut go = jit_gen_keys(rng.col,rng.rcol,rng.a,(W)(rng.b + rng.c));
                rng.a = go.a;
                rng.b = go.b;
                newkey = (unsigned)(go.keys >> 60) & 15u;   /* word 15 */
                s_keys[slot] = go.keys;
                STORE_ABC(rng,slot);
                st_n[slot] = (W)15;
            }
        }
        else if (act)
        {
            const unsigned k = q;
            const unsigned chain = s_chain[slot];
            T p[JD], pf[JD];
            T c[JRC], cf[JRC];
#pragma unroll
            for (int i = 0; i < JD; ++i)
                p[i] = sp[i*JNS + slot];
            RngT<T> rng;
            rng.col = rng_base + slot;
            rng.rcol = rsl_base + slot;
            rng.cnt = (int)st_n[slot];
            bool abc = false;           /* a, b, c are loaded (and must be stored back) */
            if (JANY_RNG)
            {
                LOAD_ABC(rng,slot);
                abc = true;
            }
            unsigned long long keys = s_keys[slot];
            W c0 = 0;                   /* randc when a, b, c were loaded: gen() increments it */
            if (abc)
                c0 = rng.c;
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
                /* _render_batch body, buffer_renderer.hpp:171-229 */
                bool bad = false;
#pragma unroll
                for (int i = 0; i < JD; ++i)
                    bad |= bad_value(p[i]);
                if (bad) /* :175-186 */
                {
#if JCOLD_BAD
                    Pt<T,JD> pb;
#pragma unroll
                    for (int i = 0; i < JD; ++i)
                        pb.v[i] = p[i];
                    const bool over = jit_record_bad(prm.stats,prm.bv_limit,jit_json_id(k),pb);
#else
                    u64 idx = atomicAdd(&prm.stats->n_bad,1ULL);
                    if (idx < FFR_MAX_BAD_RECORDED)
                    {
                        prm.stats->bad_xf[idx] = jit_json_id(k);
#pragma unroll
                        for (int i = 0; i < JD; ++i)
                            prm.stats->bad_pt[idx][i] = (double)p[i];
                    }
                    const bool over = idx + 1 > prm.bv_limit;
                    if (over)
                        atomicOr((unsigned int*)&prm.stats->abort,1u);
#endif
                    if (over)
                        gone = true;
                    else
                    {
                        /* iter.init() on the slot's own stream; pf, cf stay stale (Q3) */
                        if (!abc)
                        {
                            LOAD_ABC(rng,slot);
                            abc = true;
                            c0 = rng.c;
                        }
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
                    /* extremes of p (:188-194): after the first samples a new extreme is rare, so one
                       combined test and a branch instead of 4*D selects */
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
                    /* inclusive bounds on pf (render_iterator.hpp:72-79); NaN is out (Q4) */
                    if (jit_inb(pf))
                    {
                        ++n_plot;
                        const u64 bi = jit_index(pf); /* :202-209 */
                        if (MODES && trace)
                            trace[(u64)it*prm.chain_count + chain] = bi;
                        W *cell = buffer + bi*(1 + JR);
                        if (MODES && warp_agg)
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
            /* next iteration of this chain: colours are drawn after the last settle iteration
               (render_iterator.hpp:58-59), then the xform selection */
            const int nit = it + 1;
            if (JR > 0 && nit == 0)
            {
                if (!abc)
                {
                    LOAD_ABC(rng,slot);
                    abc = true;
                    c0 = rng.c;
                }
#pragma unroll
                for (int i = 0; i < JR; ++i)
                    c[i] = rng.num();
            }
            const int len = (chain + 1u == chain_count) ? last_len : chain_len;
            if (gone || nit >= len)
                fresh = true;          /* chain finished: this slot takes the next one below */
            else
            {
                if (abc && rng.c != c0)
                {
                    /* a draw above ran gen() inline: the block of words changed */
                    keys = jit_keys_from_rsl(rng.rcol);
                    s_keys[slot] = keys;
                }
                if (rng.cnt > 0)
                {
                    --rng.cnt;
                    newkey = (unsigned)(keys >> (4*rng.cnt)) & 15u;
                }
                else
                    newkey = JQ_GEN;   /* a converged warp runs gen() for 32 such slots */
                if (abc)
                    STORE_ABC(rng,slot);
                st_n[slot] = (W)rng.cnt;
                s_it[slot] = nit;
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
        }
        JIT_START_AND_PUSH(fresh,slot,newkey);
    }
#undef LOAD_ABC
#undef STORE_ABC

    /* merge statistics, buffer_renderer.hpp:232-246 */
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
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
        if (n_plot) atomicAdd(&prm.stats->s_plot,n_plot);
#pragma unroll
        for (int i = 0; i < JD; ++i)
        {
            atomicMin(&prm.stats->pt_min[i],f64_to_ordered((double)pmin[i]));
            atomicMax(&prm.stats->pt_max[i],f64_to_ordered((double)pmax[i]));
        }
    }
    if (lane < JNX)
        s_wxf[(tid >> 5)*JNX + lane] = wcnt;
    __syncthreads();
    if (tid < JNX)
    {
        u64 n = 0;
        for (int w = 0; w < JTPB/32; ++w)
            n += s_wxf[w*JNX + tid];
        if (n)
        {
            atomicAdd(&prm.stats->xf_dist[tid],n);
            atomicAdd(&prm.stats->s_iter,n);
        }
    }
}

/* one entry point per compilation: the render kernel proper, or -- compiled on demand, the first
   time a scatter diagnostic or the attractor-replay roofline asks for it (JIT_MODES_ENTRY) -- the
   variant that honours prm.scatter_mode */
#ifndef JIT_MODES_ENTRY
extern "C" __global__ void __launch_bounds__(JTPB,JMINB) ffr_jit_render(const RenderParams prm)
{
    jit_render_async<false>(prm);
}
#else
extern "C" __global__ void __launch_bounds__(JTPB,JMINB) ffr_jit_render_modes(const RenderParams prm)
{
    jit_render_async<true>(prm);
}
#endif
