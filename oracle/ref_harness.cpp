/*
oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A thin extern "C" harness around the UNMODIFIED reference headers, compiled from
where they lie under /root/reference/src (see oracle/Makefile). Nothing of the
reference is copied into this repo: this file only #includes the reference's
public headers and calls its public API. The output (oracle/_ref/libffr_ref.so)
is git-ignored and travels to the GPU box as a prebuilt file.

What it exposes (all through ctypes from tests/ and bench.py's reference arm):
  ref_isaac_words      rng::setSeed(u64) + rng::nextWord()      (flame_rng.hpp:47-58)
  ref_flame_info       xform order / cumulative weights / mults (flame.hpp:231-259,
                                                                 buffer_renderer.hpp:566-573)
  ref_render_chains    per chain k: rng::setSeed(seed_k); renderSeeded(L,L,bv)
                                                                (buffer_renderer.hpp:349-372)
  ref_render_mt        BufferRenderer::render(N,threads,batch,bv) timed with a
                       monotonic clock -- the CPU baseline      (buffer_renderer.hpp:269-338)
  ref_iterate_points   RenderIterator-style single steps for per-variation checks
                                                                (xform.hpp:211-227)
*/

#include "renderers/buffer_renderer.hpp"

#include <chrono>
#include <cstdint>
#include <cstring>
#include <sstream>
#include <string>

using namespace tkoz::flame;

namespace
{

thread_local std::string g_err;

inline u64 splitmix64(u64 x)
{
    u64 z = x + 0x9e3779b97f4a7c15uLL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9uLL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebuLL;
    return z ^ (z >> 31);
}

struct RefStats
{
    u64 s_iter;
    u64 s_plot;
    u64 xf_dist[64];
    double pt_min[3];
    double pt_max[3];
    u64 n_bad;
    u64 bad_xf[1024];
    double bad_pt[1024][3];
};

template <size_t dims>
int info(const Json& j, u64 *n_xf, u64 *ids, double *cw, double *mult_d,
        u64 *mult_i, u64 *cells, u64 *cell_size)
{
    Flame<dims> flame(j);
    BufferRenderer<dims> r(flame);
    *n_xf = flame.getXForms().size();
    for (size_t i = 0; i < flame.getXForms().size(); ++i)
    {
        ids[i] = flame.getXForms()[i].getID();
        cw[i] = flame.getCumulativeWeights()[i];
    }
    for (size_t i = 0; i < dims; ++i)
    {
        mult_d[i] = r.getDimMults()[i];
        mult_i[i] = r.getIndexMults()[i];
    }
    *cells = r.getBufferNumCells();
    *cell_size = r.getBufferCellSize();
    return 0;
}

template <size_t dims>
int render_chains(const Json& j, u64 base_seed, u64 chain_first,
        u64 chain_count, u64 chain_len, u64 last_len, u64 bv_limit,
        void *out, u64 out_bytes, RefStats *st)
{
    Flame<dims> flame(j);
    BufferRenderer<dims> r(flame);
    u64 bytes = r.getBuffer().size()*sizeof(hist_t);
    if (out && out_bytes != bytes)
    {
        g_err = "output size mismatch";
        return -1;
    }
    int ret = 0;
    u64 n_bad = 0;
    for (u64 k = 0; k < chain_count; ++k)
    {
        u64 len = (k+1 == chain_count && last_len) ? last_len : chain_len;
        rng::setSeed((u64)splitmix64(base_seed + chain_first + k));
        bool ok = r.renderSeeded(len,len,bv_limit);
        // renderSeeded clears the bad value lists on entry, collect per chain
        const auto& bx = r.getBadValueXForms();
        const auto& bp = r.getBadValuePoints();
        for (size_t i = 0; i < bx.size(); ++i)
        {
            if (st && n_bad < 1024)
            {
                st->bad_xf[n_bad] = bx[i];
                for (size_t d = 0; d < dims; ++d)
                    st->bad_pt[n_bad][d] = bp[i][d];
            }
            ++n_bad;
        }
        if (!ok)
        {
            ret = 1;
            break;
        }
    }
    if (st)
    {
        st->s_iter = r.getSamplesIterated();
        st->s_plot = r.getSamplesPlotted();
        for (size_t i = 0; i < r.getXFormDistribution().size() && i < 64; ++i)
            st->xf_dist[i] = r.getXFormDistribution()[i];
        for (size_t d = 0; d < dims; ++d)
        {
            st->pt_min[d] = r.getPointExtremes()[d].first;
            st->pt_max[d] = r.getPointExtremes()[d].second;
        }
        st->n_bad = n_bad;
    }
    if (out)
        memcpy(out,r.getBuffer().data(),bytes);
    return ret;
}

// the survey's pin mode: one raw u64 seed, one stream across all batches
template <size_t dims>
int render_raw(const Json& j, u64 raw_seed, u64 samples, u64 batch,
        u64 bv_limit, void *out, u64 out_bytes, RefStats *st)
{
    Flame<dims> flame(j);
    BufferRenderer<dims> r(flame);
    u64 bytes = r.getBuffer().size()*sizeof(hist_t);
    if (out && out_bytes != bytes)
    {
        g_err = "output size mismatch";
        return -1;
    }
    rng::setSeed((u64)raw_seed);
    bool ok = r.renderSeeded(samples,batch,bv_limit);
    if (st)
    {
        st->s_iter = r.getSamplesIterated();
        st->s_plot = r.getSamplesPlotted();
        for (size_t i = 0; i < r.getXFormDistribution().size() && i < 64; ++i)
            st->xf_dist[i] = r.getXFormDistribution()[i];
        for (size_t d = 0; d < dims; ++d)
        {
            st->pt_min[d] = r.getPointExtremes()[d].first;
            st->pt_max[d] = r.getPointExtremes()[d].second;
        }
        st->n_bad = r.getBadValueXForms().size();
    }
    if (out)
        memcpy(out,r.getBuffer().data(),bytes);
    return ok ? 0 : 1;
}

template <size_t dims>
int render_mt(const Json& j, u64 samples, u64 threads, u64 batch,
        u64 bv_limit, double *secs, void *out, u64 out_bytes, RefStats *st)
{
    Flame<dims> flame(j);
    BufferRenderer<dims> r(flame);
    u64 bytes = r.getBuffer().size()*sizeof(hist_t);
    if (out && out_bytes != bytes)
    {
        g_err = "output size mismatch";
        return -1;
    }
    auto t1 = std::chrono::steady_clock::now();
    bool ok = r.render(samples,threads,batch,bv_limit);
    auto t2 = std::chrono::steady_clock::now();
    *secs = std::chrono::duration<double>(t2-t1).count();
    if (st)
    {
        st->s_iter = r.getSamplesIterated();
        st->s_plot = r.getSamplesPlotted();
        for (size_t i = 0; i < r.getXFormDistribution().size() && i < 64; ++i)
            st->xf_dist[i] = r.getXFormDistribution()[i];
        for (size_t d = 0; d < dims; ++d)
        {
            st->pt_min[d] = r.getPointExtremes()[d].first;
            st->pt_max[d] = r.getPointExtremes()[d].second;
        }
        st->n_bad = r.getBadValueXForms().size();
    }
    if (out)
        memcpy(out,r.getBuffer().data(),bytes);
    return ok ? 0 : 1;
}

// one step of every point through one xform, rng seeded per point
// xf_index: index into flame.getXForms() (sorted order), -1 = final xform
template <size_t dims>
int iterate_points(const Json& j, int64_t xf_index, u64 n, const u64 *seeds,
        const double *pts_in, double *pts_out)
{
    Flame<dims> flame(j);
    const XForm<dims> *xf;
    if (xf_index < 0)
    {
        if (!flame.hasFinalXForm())
        {
            g_err = "no final xform";
            return -1;
        }
        xf = &flame.getFinalXForm();
    }
    else
    {
        if ((size_t)xf_index >= flame.getXForms().size())
        {
            g_err = "xform index out of range";
            return -1;
        }
        xf = &flame.getXForms()[xf_index];
    }
    for (u64 i = 0; i < n; ++i)
    {
        rng::setSeed((u64)seeds[i]);
        num_t tmp[dims];
        for (size_t d = 0; d < dims; ++d)
            tmp[d] = (num_t)pts_in[dims*i+d]; // the float build takes float points
        Point<num_t,dims> p(tmp);
        Point<num_t,dims> q = xf->applyIteration(p);
        for (size_t d = 0; d < dims; ++d)
            pts_out[dims*i+d] = q[d];
    }
    return 0;
}

} // namespace

#define DISPATCH(DIMS,CALL) \
    switch (DIMS) { \
    case 1: return CALL<1> ARGS; \
    case 2: return CALL<2> ARGS; \
    case 3: return CALL<3> ARGS; \
    default: g_err = "dimensions not supported"; return -1; }

extern "C"
{

const char *ref_last_error() { return g_err.c_str(); }

u64 ref_splitmix64(u64 x) { return splitmix64(x); }

// sizeof(num_t) == sizeof(hist_t): 8 for the shipped double/u64 build, 4 for the float/u32
// configuration (types.hpp:24-41)
int ref_elem_size() { return (int)sizeof(hist_t); }

void ref_isaac_words(u64 seed, u64 n, u64 *out)
{
    rng::setSeed((u64)seed);
    for (u64 i = 0; i < n; ++i)
        out[i] = rng::nextWord();
}

// randNum stream (flame_rng.hpp:67-87)
void ref_rand_nums(u64 seed, u64 n, double *out)
{
    rng::setSeed((u64)seed);
    for (u64 i = 0; i < n; ++i)
        out[i] = rng::randNum();
}

int ref_json_dims(const char *text)
{
    try
    {
        Json j{std::string(text)};
        return (int)j["dimensions"].intValue();
    }
    catch (std::exception& e)
    {
        g_err = e.what();
        return -1;
    }
}

int ref_flame_info(const char *text, u64 *n_xf, u64 *ids, double *cw,
        double *mult_d, u64 *mult_i, u64 *cells, u64 *cell_size)
{
    try
    {
        Json j{std::string(text)};
        int dims = (int)j["dimensions"].intValue();
#define ARGS (j,n_xf,ids,cw,mult_d,mult_i,cells,cell_size)
        DISPATCH(dims,info)
#undef ARGS
    }
    catch (std::exception& e)
    {
        g_err = e.what();
        return -2;
    }
}

int ref_render_chains(const char *text, u64 base_seed, u64 chain_first,
        u64 chain_count, u64 chain_len, u64 last_len, u64 bv_limit,
        void *out, u64 out_bytes, RefStats *st)
{
    try
    {
        Json j{std::string(text)};
        int dims = (int)j["dimensions"].intValue();
#define ARGS (j,base_seed,chain_first,chain_count,chain_len,last_len,bv_limit,out,out_bytes,st)
        DISPATCH(dims,render_chains)
#undef ARGS
    }
    catch (std::exception& e)
    {
        g_err = e.what();
        return -2;
    }
}

int ref_render_raw(const char *text, u64 raw_seed, u64 samples, u64 batch,
        u64 bv_limit, void *out, u64 out_bytes, RefStats *st)
{
    try
    {
        Json j{std::string(text)};
        int dims = (int)j["dimensions"].intValue();
#define ARGS (j,raw_seed,samples,batch,bv_limit,out,out_bytes,st)
        DISPATCH(dims,render_raw)
#undef ARGS
    }
    catch (std::exception& e)
    {
        g_err = e.what();
        return -2;
    }
}

int ref_render_mt(const char *text, u64 samples, u64 threads, u64 batch,
        u64 bv_limit, double *secs, void *out, u64 out_bytes, RefStats *st)
{
    try
    {
        Json j{std::string(text)};
        int dims = (int)j["dimensions"].intValue();
#define ARGS (j,samples,threads,batch,bv_limit,secs,out,out_bytes,st)
        DISPATCH(dims,render_mt)
#undef ARGS
    }
    catch (std::exception& e)
    {
        g_err = e.what();
        return -2;
    }
}

int ref_iterate_points(const char *text, int64_t xf_index, u64 n,
        const u64 *seeds, const double *pts_in, double *pts_out)
{
    try
    {
        Json j{std::string(text)};
        int dims = (int)j["dimensions"].intValue();
#define ARGS (j,xf_index,n,seeds,pts_in,pts_out)
        DISPATCH(dims,iterate_points)
#undef ARGS
    }
    catch (std::exception& e)
    {
        g_err = e.what();
        return -2;
    }
}

} // extern "C"
