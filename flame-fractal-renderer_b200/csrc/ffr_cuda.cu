/*
ffr_cuda.cu -- C ABI of libffr_cuda (include/ffr_cuda.h): context, flame blob packing,
kernel dispatch, multi-GPU sharding and peer-memory reduce. No CPU fallback: every entry
point that computes needs an sm_100 device and fails loudly without one.
*/

#include <cuda_runtime.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "ffr_kernels.cuh"
#include "ffr_jit_host.cuh"

#define FFR_VERSION_STRING "ffr-b200 0.2 (sm_100a, double/u64 + float/u32)"

namespace
{

/* polar quantities each 2-d variation reads (see calc2d in ffr_device.cuh) */
uint32_t op_need(uint32_t op)
{
    const uint32_t R2 = NEED_R2, R = NEED_R2|NEED_R, A = NEED_ANG, SC = NEED_R2|NEED_R|NEED_SC;
    switch (op)
    {
    case FFR_VAR_SWIRL: return R2;
    case FFR_VAR_HORSESHOE: return R;
    case FFR_VAR_POLAR: return A|R;
    case FFR_VAR_POLAR2: return A|R2;
    case FFR_VAR_HANDKERCHIEF: return SC; /* angle-sum form: needs y/r, x/r, not atan2 */
    case FFR_VAR_HEART: return A|R;
    case FFR_VAR_DISC: return A|R;
    case FFR_VAR_DISC2: return A;
    case FFR_VAR_FAN: return A|R;
    case FFR_VAR_RINGS: return SC;
    case FFR_VAR_SPIRAL: return SC;
    case FFR_VAR_HYPERBOLIC: return SC;
    case FFR_VAR_DIAMOND: return SC;
    case FFR_VAR_EX: return SC;
    case FFR_VAR_JULIA: return A|R;
    case FFR_VAR_POWER: return SC;
    case FFR_VAR_BLOB: return SC|A;
    case FFR_VAR_JULIAN: return A|R2;
    case FFR_VAR_JULIASCOPE: return A|R2;
    case FFR_VAR_RADIAL_BLUR: return A|R;
    case FFR_VAR_NGON: return A|R2;
    case FFR_VAR_RAYS: return R2;
    case FFR_VAR_BLADE: return R;
    case FFR_VAR_SECANT: return R;
    case FFR_VAR_TWINTRIAN: return R;
    case FFR_VAR_LOG: return A|R2;
    case FFR_VAR_SCRY: return R;
    case FFR_VAR_WEDGE: return A|R;
    case FFR_VAR_WEDGE_JULIA: return A|R2;
    case FFR_VAR_WEDGE_SPH: return A|R;
    case FFR_VAR_WHORL: return A|R;
    case FFR_VAR_SUPERSHAPE: return A|R;
    case FFR_VAR_FLOWER: return A|R;
    case FFR_VAR_CONIC: return R;
    case FFR_VAR_PARABOLA: return R;
    case FFR_VAR_BIPOLAR: return R2;
    case FFR_VAR_CPOW: return A|R2;
    case FFR_VAR_EDISC: return R2;
    case FFR_VAR_ELLIPTIC: return R2;
    case FFR_VAR_ESCHER: return A|R2;
    case FFR_VAR_LOONIE: return R2;
    default: return 0;
    }
}

/* seed-independent randmem: Isaac<u64,4>::init(flag=false), isaac.hpp:102-117 */
void isaac_m0(u64 m[16])
{
    u64 a,b,c,d,e,f,g,h;
    a = b = c = d = e = f = g = h = 0x9e3779b97f4a7c13ULL;
#define MIX() do { \
    a -= e; f ^= h >>  9; h += a; \
    b -= f; g ^= a <<  9; a += b; \
    c -= g; h ^= b >> 23; b += c; \
    d -= h; a ^= c << 15; c += d; \
    e -= a; b ^= d >> 14; d += e; \
    f -= b; c ^= e << 20; e += f; \
    g -= c; d ^= f >> 17; f += g; \
    h -= d; e ^= g << 14; g += h; } while (0)
    MIX(); MIX(); MIX(); MIX();
    for (int i = 0; i < 16; i += 8)
    {
        MIX();
        m[i+0] = a; m[i+1] = b; m[i+2] = c; m[i+3] = d;
        m[i+4] = e; m[i+5] = f; m[i+6] = g; m[i+7] = h;
    }
#undef MIX
}

/* same for Isaac<u32,4>: mix (u32) isaac.hpp:158-169, golden ratio 0x9e3779b9 */
void isaac_m0_32(unsigned int m[16])
{
    unsigned int a,b,c,d,e,f,g,h;
    a = b = c = d = e = f = g = h = 0x9e3779b9u;
#define MIX32() do { \
    a ^= b << 11; d += a; b += c; \
    b ^= c >>  2; e += b; c += d; \
    c ^= d <<  8; f += c; d += e; \
    d ^= e >> 16; g += d; e += f; \
    e ^= f << 10; h += e; f += g; \
    f ^= g >>  4; a += f; g += h; \
    g ^= h <<  8; b += g; h += a; \
    h ^= a >>  9; c += h; a += b; } while (0)
    MIX32(); MIX32(); MIX32(); MIX32();
    for (int i = 0; i < 16; i += 8)
    {
        MIX32();
        m[i+0] = a; m[i+1] = b; m[i+2] = c; m[i+3] = d;
        m[i+4] = e; m[i+5] = f; m[i+6] = g; m[i+7] = h;
    }
#undef MIX32
}

struct DeviceState
{
    int dev = -1;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    void *buffer = nullptr;        /* cells x (1+r) elements of ctx->elem bytes */
    bool own_buffer = false;
    void *d_blob = nullptr;
    void *d_colors = nullptr;
    DevStats *d_stats = nullptr;
    unsigned int *d_counter = nullptr;
    u64 *d_scratch = nullptr;      /* 2 x u64 for histogram sum/max */
    void *d_rsl = nullptr;         /* K1b randrsl scratch */
    u64 *d_trace = nullptr;        /* only while ffr_cuda_atomic_roofline records a trace */
    void *d_stage = nullptr;       /* staging for add_buffer, kept between calls */
    size_t stage_elems = 0;
    cudaStream_t copy_stream = nullptr;   /* add_buffer_async: the upload runs on a copy engine, */
    cudaEvent_t ev_copied = nullptr, ev_added = nullptr;   /* ordered against the add kernel by events */
    bool stage_busy = false;
    int sm_count = 0;
    int blocks_per_sm = 0;
    bool dirty = false;            /* holds samples not yet reduced into device 0 */
    jit::Module jmod;              /* the flame-specialised kernel loaded on this device */
    jit::Module jmod_modes;        /* K1d/K1e: its diagnostics variant (prm.scatter_mode), loaded on demand */
    int jit_blocks_per_sm = 0;
    void *d_rsl_jit = nullptr;     /* K1c randrsl scratch */
    void *d_acc = nullptr;         /* K1e accumulation tile (scrambled cell order), cells words */
    unsigned int *d_dir = nullptr; /* K1e compact tile: row directory + allocation counter (last entry) */
};

typedef void (*render_fn)(const RenderParams);

} // namespace

struct ffr_ctx
{
    uint32_t dims = 0, r = 0, cellsz = 1;
    uint32_t elem = 8;             /* sizeof(num_t) == sizeof(hist_t): 8 double/u64, 4 float/u32 */
    uint32_t size0 = 1, size1 = 1;
    u64 sizes[3] = {1,1,1};
    u64 cells = 0;
    size_t bytes = 0;
    uint32_t num_xforms = 0, num_ids = 0;
    bool has_final = false, affine_only = false, regroup = false;
    uint32_t distinct_oplists = 1;
    double divergence_ratio = 1.0;
    std::vector<unsigned char> blob;
    std::vector<unsigned char> colors;   /* T[]: xform colours in the build's precision */
    std::vector<u64> json_ids;     /* sorted index -> JSON id */
    std::vector<DeviceState> devs;
    ffr_options opt;
    uint32_t scatter_mode = FFR_SCATTER_GLOBAL;
    render_fn kernel = nullptr;
    size_t smem_bytes = 0;
    u64 launches = 0;
    std::string err;
    bool peer_enabled = false, peer_probed = false;
    /* K1c, the run-time compiled flame-specialised kernel (ffr_jit_kernel.cuh) */
    uint32_t jit_mode = 0;         /* 0 auto (lazy, large renders), 1 off, 2 on at create */
    bool jit_eligible = false, jit_ready = false, jit_failed = false, jit_cached = false;
    bool jit_cache_probed = false;   /* auto mode looked for this flame's cubin in the cache already */
    jit::Config jit_cfg;
    size_t jit_smem = 0;
    double jit_compile_s = 0.0;
    std::string jit_source, jit_err, jit_note;
    std::vector<char> jit_cubin;
    std::vector<char> jit_cubin_modes;   /* compiled by the first launch that needs it */
};

namespace
{

thread_local std::string g_create_err;

bool cuda_ok(ffr_ctx *ctx, cudaError_t e, const char *what)
{
    if (e == cudaSuccess)
        return true;
    std::string msg = std::string(what) + ": " + cudaGetErrorString(e);
    if (ctx)
        ctx->err = msg;
    else
        g_create_err = msg;
    return false;
}

#define CK(call) do { if (!cuda_ok(ctx,(call),#call)) return FFR_E_CUDA; } while (0)

/* device temporaries of one ABI call: freed on every return path */
struct DevTemp
{
    void *p = nullptr;
    DevTemp() {}
    DevTemp(const DevTemp&) = delete;
    DevTemp &operator=(const DevTemp&) = delete;
    ~DevTemp() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p,bytes); }
    template <typename U> U *as() const { return (U*)p; }
};

struct EventPair
{
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ~EventPair()
    {
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
    }
};

template <typename T, int D, int RCAP>
render_fn pick_affine(bool affine_only)
{
    return affine_only ? (render_fn)render_kernel<T,D,RCAP,true> : (render_fn)render_kernel<T,D,RCAP,false>;
}

template <typename T, int D>
render_fn pick_rcap(uint32_t r, bool affine_only)
{
    if (r == 0) return pick_affine<T,D,0>(affine_only);
    if (r <= 4) return pick_affine<T,D,4>(affine_only);
    return pick_affine<T,D,FFR_MAX_COLOR_DIMS>(affine_only);
}

/* regroup variant: colour dims <= 4 only */
template <typename T>
render_fn pick_regroup(uint32_t dims, uint32_t r, size_t *extra_smem)
{
    const int rc = (r == 0) ? 0 : 4;
    *extra_smem = FFR_SMEM_REGROUP_BYTES(dims,rc,sizeof(T));
    switch (dims*10 + rc)
    {
    case 10: return (render_fn)render_kernel_regroup<T,1,0>;
    case 14: return (render_fn)render_kernel_regroup<T,1,4>;
    case 20: return (render_fn)render_kernel_regroup<T,2,0>;
    case 24: return (render_fn)render_kernel_regroup<T,2,4>;
    case 30: return (render_fn)render_kernel_regroup<T,3,0>;
    case 34: return (render_fn)render_kernel_regroup<T,3,4>;
    default: return nullptr;
    }
}

template <typename T>
render_fn pick_kernel(uint32_t dims, uint32_t r, bool affine_only)
{
    switch (dims)
    {
    case 1: return pick_rcap<T,1>(r,affine_only);
    case 2: return pick_rcap<T,2>(r,affine_only);
    case 3: return pick_rcap<T,3>(r,affine_only);
    default: return nullptr;
    }
}

/* flatten the caller's desc into the device blob (DevFlame | DevXForm[] | DevVar[]) */
template <typename T>
bool pack_blob(ffr_ctx *ctx, const ffr_flame_desc *d, std::string &err)
{
    if (!d || d->dims < 1 || d->dims > FFR_MAX_DIMS)
    {
        err = "dimensions not supported";
        return false;
    }
    if (d->elem_size != sizeof(T))
    {
        err = "elem_size must be 8 (double/u64) or 4 (float/u32)";
        return false;
    }
    if (d->color_dims > FFR_MAX_COLOR_DIMS)
    {
        err = "too many color dimensions";
        return false;
    }
    if (d->num_xforms == 0 || !d->xforms || !d->xfcw)
    {
        err = "Flame(): no xforms";
        return false;
    }
    if (d->num_xforms > FFR_MAX_XFORMS || d->num_xform_ids > FFR_MAX_XFORMS)
    {
        err = "libffr_cuda supports at most 64 xforms";
        return false;
    }
    DevFlameT<T> hdr;
    memset(&hdr,0,sizeof(hdr));
    hdr.dims = d->dims;
    hdr.r = d->color_dims;
    hdr.has_final = d->has_final && d->final_xform;
    hdr.num_xforms = d->num_xforms;
    hdr.num_ids = d->num_xform_ids;
    /* BufferRenderer::_init, buffer_renderer.hpp:114-140, in num_t arithmetic;
       scale_adjust_down_v<num_t> = 1 - emach (constants.hpp:25-30,59-62) */
    const T scale_adjust_down = sizeof(T) == 8 ? (T)(1.0 - (double)(float)(1.0 / (double)(1L << 52)))
                                               : (T)(1.0F - 1.0F / (float)(1 << 23));
    u64 cells = 1;
    for (uint32_t i = 0; i < d->dims; ++i)
    {
        if (d->size[i] == 0 || !(d->bounds_lo[i] < d->bounds_hi[i]))
        {
            err = "Flame(): bad size or bounds";
            return false;
        }
        hdr.lo[i] = (T)d->bounds_lo[i];
        hdr.hi[i] = (T)d->bounds_hi[i];
        hdr.mult_d[i] = (T)(d->size[i]) / (hdr.hi[i] - hdr.lo[i]);
        hdr.mult_d[i] *= scale_adjust_down;
        hdr.mult_i[i] = cells;
        cells *= d->size[i];
        if (cells >= (1ULL << 48))
        {
            err = "BufferRenderer(): histogram too big";
            return false;
        }
    }
    hdr.cells = cells;
    hdr.cell = 1 + d->color_dims;
    for (uint32_t i = 0; i < FFR_MAX_XFORMS; ++i)
        hdr.xfcw[i] = (i < d->num_xforms) ? (T)d->xfcw[i] : (T)2.0; /* padding never < r */

    std::vector<DevXFormT<T>> xfs;
    std::vector<DevVarT<T>> vars;
    std::vector<T> colors;
    ctx->json_ids.clear();
    bool affine_only = true, uses_rng = false;
    const uint32_t total = d->num_xforms + (hdr.has_final ? 1 : 0);
    for (uint32_t i = 0; i < total; ++i)
    {
        const ffr_xform &x = (i < d->num_xforms) ? d->xforms[i] : *d->final_xform;
        DevXFormT<T> dx;
        memset(&dx,0,sizeof(dx));
        for (int k = 0; k < 9; ++k)
        {
            dx.pre_A[k] = (T)x.pre_A[k];
            dx.post_A[k] = (T)x.post_A[k];
        }
        for (int k = 0; k < 3; ++k)
        {
            dx.pre_b[k] = (T)x.pre_b[k];
            dx.post_b[k] = (T)x.post_b[k];
        }
        dx.color_speed = (T)x.color_speed;
        dx.var_begin = (uint32_t)vars.size();
        dx.var_count = x.num_vars;
        dx.flags = (x.has_pre ? XF_HAS_PRE : 0) | (x.has_post ? XF_HAS_POST : 0);
        if (i < d->num_xforms)
        {
            if (x.id >= FFR_MAX_XFORMS)
            {
                err = "xform id out of range";
                return false;
            }
            ctx->json_ids.push_back(x.id);
            dx.json_id = (uint32_t)x.id;
        }
        else
            dx.json_id = 0xffffffffu;
        if (x.has_color && x.color && d->color_dims)
        {
            dx.flags |= XF_HAS_COLOR;
            dx.color_off = (uint32_t)colors.size();
            for (uint32_t k = 0; k < d->color_dims; ++k)
                colors.push_back((T)x.color[k]);
        }
        for (uint32_t k = 0; k < x.num_vars; ++k)
        {
            const ffr_variation &v = x.vars[k];
            if (v.op < 1 || v.op > FFR_VAR_COUNT)
            {
                err = "unknown variation opcode";
                return false;
            }
            const bool v2d = v.op >= FFR_VAR_FIRST_2D && v.op <= FFR_VAR_LAST_2D;
            if (v2d && d->dims < 2)
            {
                err = "2-d variation in a 1-d flame";
                return false;
            }
            if (v2d && d->dims > 2 && (v.axis_x >= d->dims || v.axis_y >= d->dims || v.axis_x == v.axis_y))
            {
                err = "axis index out of range";
                return false;
            }
            DevVarT<T> dv;
            memset(&dv,0,sizeof(dv));
            dv.op = v.op;
            dv.axis_x = (d->dims == 2) ? 0 : v.axis_x;
            dv.axis_y = (d->dims == 2) ? 1 : v.axis_y;
            dv.need = op_need(v.op) | (var_uses_rng(v.op) ? NEED_RNG : 0u);
            dv.weight = (T)v.weight;
            for (int q = 0; q < FFR_MAX_VAR_PARAMS; ++q)
                dv.p[q] = (T)v.params[q];
            dx.need |= dv.need;
            if (v.op != FFR_VAR_LINEAR)
                affine_only = false;
            if (var_uses_rng(v.op))
            {
                dx.flags |= XF_USES_RNG;
                uses_rng = true;
            }
            vars.push_back(dv);
        }
        xfs.push_back(dx);
    }
    {
        /* xforms whose variation opcode sequences differ make lanes diverge */
        std::vector<std::vector<uint32_t>> lists;
        for (uint32_t i = 0; i < d->num_xforms; ++i)
        {
            std::vector<uint32_t> l;
            for (uint32_t k = 0; k < xfs[i].var_count; ++k)
                l.push_back(vars[xfs[i].var_begin + k].op);
            if (std::find(lists.begin(),lists.end(),l) == lists.end())
                lists.push_back(l);
        }
        ctx->distinct_oplists = (uint32_t)lists.size();
        /* Divergence model for choosing K1b. In K1 a warp executes, per position j of the
           variation loop, every DISTINCT non-linear opcode its lanes hold there; in K1b a warp
           runs one xform, i.e. the selection-weighted mean number of non-linear variations,
           plus the per-iteration sort (~1.5 variation bodies). Measured: csci6360 (ratio 3.1)
           gains 35 % from K1b, tkoz_test3 (1.8) loses 10 %. */
        double direct = 0.0;
        size_t maxlen = 0;
        for (auto& l : lists) maxlen = std::max(maxlen,l.size());
        for (size_t j = 0; j < maxlen; ++j)
        {
            std::vector<uint32_t> seen;
            for (uint32_t i = 0; i < d->num_xforms; ++i)
                if (j < xfs[i].var_count)
                {
                    uint32_t op = vars[xfs[i].var_begin + j].op;
                    if (op != FFR_VAR_LINEAR && std::find(seen.begin(),seen.end(),op) == seen.end())
                        seen.push_back(op);
                }
            direct += (double)seen.size();
        }
        double grouped = 1.5, prev = 0.0;
        for (uint32_t i = 0; i < d->num_xforms; ++i)
        {
            uint32_t nl = 0;
            for (uint32_t k = 0; k < xfs[i].var_count; ++k)
                nl += vars[xfs[i].var_begin + k].op != FFR_VAR_LINEAR;
            grouped += (d->xfcw[i] - prev) * nl;
            prev = d->xfcw[i];
        }
        ctx->divergence_ratio = direct / grouped;
    }
    if (colors.empty())
        colors.push_back((T)0);
    ctx->colors.assign((const unsigned char*)colors.data(),
        (const unsigned char*)colors.data() + colors.size()*sizeof(T));
    hdr.num_vars = (uint32_t)vars.size();
    hdr.uses_rng = uses_rng;
    hdr.xf_off = (uint32_t)sizeof(DevFlameT<T>);
    hdr.var_off = hdr.xf_off + (uint32_t)(xfs.size()*sizeof(DevXFormT<T>));
    hdr.total_bytes = hdr.var_off + (uint32_t)(vars.size()*sizeof(DevVarT<T>));
    hdr.total_bytes = (hdr.total_bytes + 15u) & ~15u;
    if (hdr.total_bytes > 96*1024)
    {
        err = "flame too large for the shared-memory blob";
        return false;
    }
    ctx->blob.assign(hdr.total_bytes,0);
    memcpy(ctx->blob.data(),&hdr,sizeof(hdr));
    memcpy(ctx->blob.data()+hdr.xf_off,xfs.data(),xfs.size()*sizeof(DevXFormT<T>));
    if (!vars.empty())
        memcpy(ctx->blob.data()+hdr.var_off,vars.data(),vars.size()*sizeof(DevVarT<T>));
    ctx->dims = d->dims;
    for (uint32_t i = 0; i < 3; ++i)
        ctx->sizes[i] = i < d->dims ? d->size[i] : 1;
    ctx->size0 = (uint32_t)d->size[0];
    ctx->size1 = d->dims > 1 ? (uint32_t)d->size[1] : 1;
    ctx->r = d->color_dims;
    ctx->cellsz = hdr.cell;
    ctx->cells = cells;
    ctx->elem = (uint32_t)sizeof(T);
    ctx->bytes = (size_t)cells*hdr.cell*sizeof(T);
    ctx->num_xforms = d->num_xforms;
    ctx->num_ids = d->num_xform_ids;
    ctx->has_final = hdr.has_final;
    ctx->affine_only = affine_only;
    return true;
}

void init_stats_host(DevStats &s)
{
    memset(&s,0,sizeof(s));
    for (int i = 0; i < 3; ++i)
    {
        s.pt_min[i] = f64_to_ordered(INFINITY);
        s.pt_max[i] = f64_to_ordered(-INFINITY);
    }
}

int setup_device(ffr_ctx *ctx, DeviceState &ds, int dev, const ffr_options &opt)
{
    ds.dev = dev;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop,dev));
    if (prop.major < 10)
    {
        ctx->err = "device " + std::to_string(dev) + " (" + prop.name + ") is not sm_100 or newer; "
            "libffr_cuda has no fallback path";
        return FFR_E_NODEVICE;
    }
    ds.sm_count = prop.multiProcessorCount;
    if (opt.stream)
        ds.stream = (cudaStream_t)opt.stream;
    else
    {
        CK(cudaStreamCreateWithFlags(&ds.stream,cudaStreamNonBlocking));
        ds.own_stream = true;
    }
    if (opt.external_buffer)
        ds.buffer = (u64*)opt.external_buffer;
    else
    {
        CK(cudaMalloc(&ds.buffer,ctx->bytes));
        ds.own_buffer = true;
        CK(cudaMemsetAsync(ds.buffer,0,ctx->bytes,ds.stream));
    }
    CK(cudaMalloc(&ds.d_blob,ctx->blob.size()));
    CK(cudaMemcpyAsync(ds.d_blob,ctx->blob.data(),ctx->blob.size(),cudaMemcpyHostToDevice,ds.stream));
    CK(cudaMalloc(&ds.d_colors,ctx->colors.size()));
    CK(cudaMemcpyAsync(ds.d_colors,ctx->colors.data(),ctx->colors.size(),
        cudaMemcpyHostToDevice,ds.stream));
    CK(cudaMalloc(&ds.d_stats,sizeof(DevStats)));
    DevStats init;
    init_stats_host(init);
    CK(cudaMemcpyAsync(ds.d_stats,&init,sizeof(init),cudaMemcpyHostToDevice,ds.stream));
    CK(cudaMalloc(&ds.d_counter,sizeof(unsigned int)));
    CK(cudaMalloc(&ds.d_scratch,2*sizeof(u64)));
    u64 m0[16];
    isaac_m0(m0);
    CK(cudaMemcpyToSymbolAsync(c_isaac_m0,m0,sizeof(m0),0,cudaMemcpyHostToDevice,ds.stream));
    unsigned int m0_32[16];
    isaac_m0_32(m0_32);
    CK(cudaMemcpyToSymbolAsync(c_isaac_m0_32,m0_32,sizeof(m0_32),0,cudaMemcpyHostToDevice,ds.stream));
    CK(cudaFuncSetAttribute((const void*)ctx->kernel,cudaFuncAttributeMaxDynamicSharedMemorySize,
        (int)ctx->smem_bytes));
    int nb = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb,(const void*)ctx->kernel,FFR_TPB,ctx->smem_bytes));
    if (nb < 1)
    {
        ctx->err = "render kernel does not fit on the device";
        return FFR_E_CUDA;
    }
    if (opt.blocks_per_sm && (int)opt.blocks_per_sm < nb)
        nb = (int)opt.blocks_per_sm;
    ds.blocks_per_sm = nb;
    if (ctx->regroup)
        CK(cudaMalloc(&ds.d_rsl,(size_t)ds.sm_count*nb*16*FFR_TPB*ctx->elem));
    CK(cudaStreamSynchronize(ds.stream));
    return FFR_OK;
}


/* ---- K1c: flame-specialised kernel, compiled at run time (ffr_jit_host.cuh) ---- */

int env_int(const char *name, int dflt)
{
    const char *e = getenv(name);
    return (e && *e) ? atoi(e) : dflt;
}

/* K1c keeps one slot queue per xform and colour state in registers */
bool jit_supported(const ffr_ctx *ctx)
{
    return ctx->num_xforms >= 1 && ctx->num_xforms <= 8 && ctx->r <= 4;
}

/* generate the source for this flame and compile it (no device needed) */
bool jit_prepare(ffr_ctx *ctx, bool cache_only = false)
{
    if (!ctx->jit_cubin.empty())
        return true;
    jit::Config &cfg = ctx->jit_cfg;
    const bool uses_rng = ctx->elem == 8 ? ((const DevFlameT<double>*)ctx->blob.data())->uses_rng != 0
                                         : ((const DevFlameT<float>*)ctx->blob.data())->uses_rng != 0;
    u64 m0[16];
    unsigned int m0_32[16];
    isaac_m0(m0);
    isaac_m0_32(m0_32);
    /* K1e: pure-affine flames get the kernel written for them (ffr_jit_affine.cuh) */
    if (ctx->affine_only && env_int("FFR_JIT_AFFINE",1) != 0)
    {
        cfg.affine = true;
        cfg.async = false;
        cfg.tpb = env_int("FFR_JIT_TPB",256);
        cfg.minb = env_int("FFR_JIT_MINB",3);
        cfg.inline_math = false;
        cfg.ns = cfg.tpb;
        cfg.cap = cfg.tpb;
        /* Accumulation tile in scrambled cell order (fold_acc_kernel) for power-of-two cell counts.
           Measured on B200 (samples/s, linear -> scrambled): sierpinski@1024^2 (8 MiB) 1.40e11 ->
           2.04e11, barnsley@2048^2 (32 MiB) 1.57e11 -> 1.88e11 with single cells scrambled; at
           128 MiB single cells lose for a dense attractor (barnsley@4096^2 1.81e11 -> 1.68e11: every
           touched cell becomes a sector of its own and the touched sectors no longer fit L2) while
           scrambling whole 32-byte sectors still wins (1.86e11; sierpinski@4096^2 1.84e11 -> 2.03e11,
           sierpinski_3d@256^3 1.63e11 -> 1.76e11); at 1 GiB (sierpinski_3d@512^3) every variant loses
           (8.2e10 -> 5.4e10), so large buffers are scattered into directly. */
        cfg.acc_mul = 0;
        cfg.acc_gran = 0;
        {
            const u64 bytes = ctx->cells*ctx->elem;
            const u64 max_bytes = (u64)env_int("FFR_ACC_MAX_MB",128) << 20;
            if (env_int("FFR_K1E_SCRAMBLE",1) != 0 && ctx->r == 0 && ctx->cells >= 4096 &&
                    (ctx->cells & (ctx->cells - 1)) == 0 && bytes <= max_bytes && ctx->cells <= (1ULL << 31))
            {
                cfg.acc_mul = 0x9E3779B1u;
                /* cells per 32-byte sector stay together above 32 MiB */
                const int sector = ctx->elem == 8 ? 2 : 3;
                cfg.acc_gran = (unsigned)env_int("FFR_ACC_GRAN",bytes <= (32u << 20) ? 0 : sector);
            }
        }
        /* Compact tile behind a row directory for buffers beyond the TLB reach (fold_dir_kernel):
           sierpinski_3d@512^3 (1 GiB) 8.2e10 -> see DESIGN.md. Tile <= 192 MiB. */
        cfg.dir_cap = 0;
        {
            const u64 bytes = ctx->cells*ctx->elem;
            const u64 min_bytes = (u64)env_int("FFR_DIR_MIN_MB",257) << 20;
            if (!cfg.acc_mul && env_int("FFR_K1E_DIR",1) != 0 && ctx->r == 0 && bytes >= min_bytes &&
                    ctx->cells % (1u << FFR_DIR_ROW_SHIFT) == 0 && (ctx->cells >> FFR_DIR_ROW_SHIFT) < 0xfffffff0ULL)
            {
                const u64 tile_bytes = (u64)env_int("FFR_DIR_TILE_MB",192) << 20;
                cfg.dir_cap = (unsigned)std::min<u64>(tile_bytes/((u64)ctx->elem << FFR_DIR_ROW_SHIFT),
                                                      ctx->cells >> FFR_DIR_ROW_SHIFT);
                /* rows as BLOCKS of cells where every size is a multiple of the block's extent
                   (FFR_DIR_BLOCKED=0: 512 consecutive cells, the round-1 layout) */
                static const unsigned shape[4][3] = {{0,0,0},{9,0,0},{5,4,0},{3,3,3}};
                bool ok = env_int("FFR_DIR_BLOCKED",1) != 0 && ctx->dims >= 1 && ctx->dims <= 3;
                for (uint32_t d = 0; ok && d < ctx->dims; ++d)
                    ok = ctx->sizes[d] % (1u << shape[ctx->dims][d]) == 0 && ctx->sizes[d] < (1ULL << 31);
                for (uint32_t d = 0; d < 3; ++d)
                    cfg.dir_blk[d] = ok ? shape[ctx->dims][d] : 0u;
                /* Directory cache in shared memory: ONE block of 768 chains per SM instead of three of
                   256 (same 24 warps, same ISAAC footprint), which leaves ~35 KB for a 2-way cache of
                   tag|slot words. Needs slots < 0xffff and rows < 0xffff sets-blocks (both hold for
                   every buffer up to 128 GiB). FFR_DIR_CACHE=0: off. */
                if (env_int("FFR_DIR_CACHE",1) != 0 && !getenv("FFR_JIT_TPB") && !getenv("FFR_JIT_MINB") &&
                        cfg.dir_cap < 0xffffu)
                {
                    const size_t isaac = (size_t)32*768*ctx->elem;
                    const size_t room = (size_t)227*1024 - 2048;
                    unsigned sets = 0;
                    if (isaac + 1024 < room)
                        for (sets = 8192; sets >= 512 && isaac + 1024 + (size_t)sets*8 > room; sets >>= 1) {}
                    if (sets >= 512 && (ctx->cells >> FFR_DIR_ROW_SHIFT) < (u64)0xffffu*sets)
                    {
                        cfg.dir_cache_sets = sets;
                        cfg.tpb = 768;
                        cfg.minb = 1;
                        cfg.ns = cfg.cap = cfg.tpb;
                    }
                }
            }
        }
        std::string why;
        if (cfg.tpb >= 32 && cfg.tpb <= 1024 && cfg.tpb % 32 == 0)
            ctx->jit_source = ctx->elem == 8 ? jit::generate_affine<double>(ctx->blob,m0,m0_32,cfg,&cfg.npair,why)
                                             : jit::generate_affine<float>(ctx->blob,m0,m0_32,cfg,&cfg.npair,why);
        if (!ctx->jit_source.empty())
        {
            /* randmem + randrsl columns, then the coefficient table [pair][xform] */
            ctx->jit_smem = (size_t)32*cfg.tpb*ctx->elem + (size_t)cfg.npair*ctx->num_xforms*2*ctx->elem +
                (size_t)cfg.dir_cache_sets*8;
            long spills = 0;
            double secs = 0.0;
            if (!jit::compile(ctx->jit_source,ctx->jit_cubin,ctx->jit_err,&secs,&ctx->jit_cached,&spills,cache_only))
            {
                ctx->jit_cubin.clear();
                return false;
            }
            ctx->jit_compile_s += secs;
            ctx->jit_note += std::string("K1e pure-affine kernel, ") +
                (cfg.acc_mul ? (cfg.acc_gran ? "sector-scrambled accumulation tile, " : "cell-scrambled accumulation tile, ") : "") +
                (cfg.dir_cap ? "compact tile of " + std::to_string(cfg.dir_cap) + " rows" +
                    ((cfg.dir_blk[0] | cfg.dir_blk[1] | cfg.dir_blk[2]) ? " (blocks of " + std::to_string(1u << cfg.dir_blk[0]) +
                        (ctx->dims > 1 ? "x" + std::to_string(1u << cfg.dir_blk[1]) : std::string()) +
                        (ctx->dims > 2 ? "x" + std::to_string(1u << cfg.dir_blk[2]) : std::string()) + " cells)" : std::string()) +
                    (cfg.dir_cache_sets ? " + directory cache of " + std::to_string(2*cfg.dir_cache_sets) + " entries in shared memory" : std::string()) + ", "
                  : std::string()) +
                std::to_string(cfg.npair) + " table rows, tpb " +
                std::to_string(cfg.tpb) + ": " + std::to_string(spills) + " spill bytes; ";
            return true;
        }
        cfg.affine = false;
        ctx->jit_note += why + "; ";
    }
    cfg.async = env_int("FFR_JIT_ASYNC",1) != 0;
    /* K1d: 320 threads x 2 blocks (96 registers) for the double build -- more warps there mean an
       80-register cap, measured slower. The float build needs 72 registers and
       has slots to spare: 448 threads measured +9 % over 320 (2.99e10 -> 3.25e10 on csci6360@4096^2) */
    cfg.tpb = env_int("FFR_JIT_TPB",cfg.async ? (ctx->elem == 4 ? 448 : 320) : 256);
    cfg.minb = env_int("FFR_JIT_MINB",2);
    cfg.inline_math = env_int("FFR_JIT_INLINE_MATH",0) != 0;
    /* slots per block: as many as fit next to a second block in the 227 KB of an SM.
       Per slot: ISAAC randmem[16] + randa/b/c/cnt, the point, the colour, then K1d: 16 packed
       selections, iteration and chain numbers + one ring entry per queue; K1c: two queue entries per xform */
    const size_t per_slot = (size_t)20*ctx->elem + (size_t)(ctx->dims + ctx->r)*ctx->elem +
        (cfg.async ? 16u : 4u*ctx->num_xforms) + (cfg.async && uses_rng ? (size_t)16*ctx->elem : 0u);
    const size_t nq = ctx->num_xforms + 1u;
    auto pow2ceil = [](int v) { int p = 64; while (p < v) p *= 2; return p; };
    auto smem_for = [&](int n) { return per_slot*(size_t)n + (cfg.async ? 2u*nq*(size_t)pow2ceil(n) : 0u); };
    int ns = env_int("FFR_JIT_NS",0);
    if (ns <= 0)
    {
        const size_t budget = (227u*1024u)/(size_t)cfg.minb - 2048u;
        ns = 1024;
        while (ns > 64 && smem_for(ns) > budget)
            ns -= 32;
        if (!cfg.async)
            ns -= ns % cfg.tpb;
    }
    ns -= ns % 32;       /* whole warps seed the slots */
    if (cfg.async && !getenv("FFR_JIT_TPB"))
    {
        /* keep >= 32 slots per queue out of flight so that full chunks can always be popped */
        int t = ns - 32*(int)nq;
        t -= t % 32;
        cfg.tpb = std::max(128,std::min(cfg.tpb,t));
    }
    if (ns < 64 || ns > 32768 || cfg.tpb < 64 || cfg.tpb % 32 || cfg.tpb > 1024)
    {
        ctx->jit_err = "K1c: no valid slot count for this flame";
        return false;
    }
    cfg.ns = ns;
    cfg.cap = pow2ceil(ns);
    ctx->jit_smem = smem_for(ns);
    for (int attempt = 0; attempt < 3; ++attempt)
    {
        ctx->jit_source = ctx->elem == 8 ? jit::generate<double>(ctx->blob,ctx->colors,m0,m0_32,cfg)
                                         : jit::generate<float>(ctx->blob,ctx->colors,m0,m0_32,cfg);
        long spills = 0;
        double secs = 0.0;
        if (!jit::compile(ctx->jit_source,ctx->jit_cubin,ctx->jit_err,&secs,&ctx->jit_cached,&spills,cache_only))
        {
            ctx->jit_cubin.clear();
            return false;
        }
        ctx->jit_compile_s += secs;
        ctx->jit_note += std::string(cfg.async ? "K1d queue-scheduled kernel, " : "K1c lock-step kernel, ") +
            "tpb " + std::to_string(cfg.tpb) + ": " + std::to_string(spills) + " spill bytes; ";
        /* 320 threads x 2 blocks cap the kernel at 96 registers; a flame whose xforms spill there
           runs faster with 256 threads (128 registers) than with spills through a thrashed L1 */
        if (spills > 32 && cfg.tpb > 256 && !getenv("FFR_JIT_TPB"))
        {
            cfg.tpb = cfg.tpb > 320 ? 320 : 256;
            continue;
        }
        break;
    }
    return true;
}

/* compile if necessary and load the kernel on every device of the context */
int jit_activate(ffr_ctx *ctx)
{
    if (ctx->jit_ready)
        return FFR_OK;
    if (ctx->jit_failed || !ctx->jit_eligible)
        return FFR_E_UNSUPPORTED;
    if (!jit_prepare(ctx))
    {
        ctx->jit_failed = true;
        return FFR_E_UNSUPPORTED;
    }
    if (ctx->affine_only && !ctx->jit_cfg.affine && ctx->jit_mode == 0)
    {
        ctx->jit_failed = true;      /* auto mode: K1e does not cover this flame, K1 stays */
        ctx->jit_cubin.clear();
        return FFR_E_UNSUPPORTED;
    }
    for (DeviceState &ds : ctx->devs)
    {
        CK(cudaSetDevice(ds.dev));
        CK(cudaFree(0));
        if (!jit::load(ctx->jit_cubin,ctx->jit_smem,ds.jmod,ctx->jit_err))
        {
            ctx->jit_failed = true;
            return FFR_E_UNSUPPORTED;
        }
        int nb = 0;
        jit::Api &a = jit::api(true);
        if (a.Occupancy(&nb,ds.jmod.fn,ctx->jit_cfg.tpb,ctx->jit_smem) != CUDA_SUCCESS || nb < 1)
        {
            ctx->jit_err = "K1c does not fit on the device";
            ctx->jit_failed = true;
            return FFR_E_UNSUPPORTED;
        }
        if (ctx->opt.blocks_per_sm && (int)ctx->opt.blocks_per_sm < nb)
            nb = (int)ctx->opt.blocks_per_sm;
        ds.jit_blocks_per_sm = nb;
        CK(cudaMalloc(&ds.d_rsl_jit,(size_t)ds.sm_count*nb*16*ctx->jit_cfg.ns*ctx->elem));
        if (ctx->jit_cfg.affine && ctx->jit_cfg.acc_mul)
        {
            CK(cudaMalloc(&ds.d_acc,(size_t)ctx->cells*ctx->elem));
            CK(cudaMemsetAsync(ds.d_acc,0,(size_t)ctx->cells*ctx->elem,ds.stream));
        }
        if (ctx->jit_cfg.affine && ctx->jit_cfg.dir_cap)
        {
            const size_t tile = ((size_t)ctx->jit_cfg.dir_cap << FFR_DIR_ROW_SHIFT)*ctx->elem;
            const size_t rows = (size_t)(ctx->cells >> FFR_DIR_ROW_SHIFT);
            CK(cudaMalloc(&ds.d_acc,tile));
            CK(cudaMemsetAsync(ds.d_acc,0,tile,ds.stream));
            CK(cudaMalloc(&ds.d_dir,(rows + 1)*sizeof(unsigned int)));
            CK(cudaMemsetAsync(ds.d_dir,0xff,rows*sizeof(unsigned int),ds.stream));
            CK(cudaMemsetAsync(ds.d_dir + rows,0,sizeof(unsigned int),ds.stream));
        }
    }
    ctx->jit_ready = true;
    return FFR_OK;
}

/* auto mode: a render this large repays the compile (seconds) many times over */
void jit_maybe(ffr_ctx *ctx, u64 samples)
{
    if (ctx->jit_ready || ctx->jit_failed || !ctx->jit_eligible || ctx->jit_mode != 0)
        return;
    const char *e = getenv("FFR_JIT_MIN_SAMPLES");
    /* ~2 s of NVRTC against ~4e-11 s saved per sample over the interpreter kernels (variation
       flames); ~0.4 s against ~4.5e-12 s per sample for pure-affine flames (K1e 1.85e11/s over K1
       1.0e11/s); the cubin is cached on disk, so a repeated flame pays nothing */
    const double min_samples = (e && *e) ? atof(e) : (ctx->affine_only ? 2e11 : 5e10);
    if ((double)samples >= min_samples)
        jit_activate(ctx);
    else if (!ctx->jit_cache_probed && env_int("FFR_JIT_USE_CACHED",1) != 0)
    {
        ctx->jit_cache_probed = true;
        /* a smaller render: the compiled kernel only if this flame's cubin is already in the
           process or disk cache (an earlier run paid for it); loading it takes milliseconds */
        const std::string note = ctx->jit_note;
        if (jit_prepare(ctx,true))
            jit_activate(ctx);
        else
        {
            ctx->jit_err.clear();
            ctx->jit_note = note;
            ctx->jit_cubin.clear();
            ctx->jit_cfg = jit::Config();
        }
    }
}

/* the diagnostics variant of the run-time compiled kernel: same source, other entry point */
int jit_load_modes(ffr_ctx *ctx, DeviceState &ds)
{
    if (ds.jmod_modes.fn)
        return FFR_OK;
    if (ctx->jit_cubin_modes.empty())
    {
        double secs = 0.0;
        bool cached = false;
        long spills = 0;
        if (!jit::compile("#define JIT_MODES_ENTRY 1\n" + ctx->jit_source,ctx->jit_cubin_modes,ctx->jit_err,
                &secs,&cached,&spills,false))
        {
            ctx->jit_cubin_modes.clear();
            ctx->err = "run-time compilation of the diagnostics kernel failed: " + ctx->jit_err;
            return FFR_E_UNSUPPORTED;
        }
        ctx->jit_compile_s += secs;
    }
    if (!jit::load(ctx->jit_cubin_modes,ctx->jit_smem,ds.jmod_modes,ctx->jit_err,"ffr_jit_render_modes"))
    {
        ctx->err = ctx->jit_err;
        return FFR_E_CUDA;
    }
    return FFR_OK;
}

/* K2b / K2c: what a K1e launch (or the attractor replay) left in its accumulation tile goes into
   the buffer in the reference's cell order; the tile is all zero afterwards */
int fold_tiles(ffr_ctx *ctx, DeviceState &ds)
{
    if (!ctx->jit_ready || !ds.d_acc)
        return FFR_OK;
    if (ds.d_dir)
    {
        /* K2c: the compact tile's rows into the buffer */
        const u64 rows = ctx->cells >> FFR_DIR_ROW_SHIFT;
        const unsigned fgrid = (unsigned)std::min<u64>((rows + 7)/8,(u64)ds.sm_count*16);
        DirGeom g;
        u64 mult = 1;
        for (uint32_t d = 0; d < 3; ++d)
        {
            g.blk[d] = ctx->jit_cfg.dir_blk[d];
            g.rows[d] = (uint32_t)std::max<u64>(1,ctx->sizes[d] >> g.blk[d]);
            g.mult[d] = mult;
            mult *= ctx->sizes[d];
        }
        if (ctx->elem == 8)
            fold_dir_kernel<u64><<<fgrid,256,0,ds.stream>>>((u64*)ds.d_acc,(u64*)ds.buffer,ds.d_dir,rows,g);
        else
            fold_dir_kernel<unsigned int><<<fgrid,256,0,ds.stream>>>((unsigned int*)ds.d_acc,(unsigned int*)ds.buffer,ds.d_dir,rows,g);
    }
    else
    {
        /* K2b: the launch's scrambled tile into the buffer (reference cell order) */
        uint32_t inv = ctx->jit_cfg.acc_mul;      /* Newton: x <- x*(2 - m*x) doubles the valid bits */
        for (int i = 0; i < 5; ++i)
            inv *= 2u - ctx->jit_cfg.acc_mul*inv;
        const unsigned fgrid = (unsigned)std::min<u64>((ctx->cells + 255)/256,(u64)ds.sm_count*16);
        if (ctx->elem == 8)
            fold_acc_kernel<u64><<<fgrid,256,0,ds.stream>>>((u64*)ds.d_acc,(u64*)ds.buffer,ctx->cells,inv,ctx->jit_cfg.acc_gran);
        else
            fold_acc_kernel<unsigned int><<<fgrid,256,0,ds.stream>>>((unsigned int*)ds.d_acc,(unsigned int*)ds.buffer,ctx->cells,inv,ctx->jit_cfg.acc_gran);
    }
    ++ctx->launches;
    CK(cudaGetLastError());
    return FFR_OK;
}

int launch_render(ffr_ctx *ctx, DeviceState &ds, u64 chain_first, u64 chain_count, u64 chain_len,
        u64 last_len, u64 base_seed, u64 bv_limit)
{
    if (chain_count == 0)
        return FFR_OK;
    CK(cudaSetDevice(ds.dev));
    RenderParams prm;
    prm.blob = ds.d_blob;
    prm.colors = ds.d_colors;
    prm.buffer = ds.buffer;
    prm.stats = ds.d_stats;
    prm.work_counter = ds.d_counter;
    prm.rsl_scratch = ds.d_rsl;
    prm.trace = ds.d_trace;
    prm.chain_first = chain_first;
    prm.chain_count = chain_count;
    prm.chain_len = chain_len;
    prm.last_len = last_len;
    prm.base_seed = base_seed;
    prm.bv_limit = bv_limit;
    prm.blob_bytes = (uint32_t)ctx->blob.size();
    prm.scatter_mode = ctx->scatter_mode;
    prm.acc = ds.d_acc;
    prm.dir = ds.d_dir;
    prm.dir_next = ds.d_dir ? ds.d_dir + (ctx->cells >> FFR_DIR_ROW_SHIFT) : nullptr;
    if (chain_len >= (1ULL << 31) - 64)
    {
        ctx->err = "chain length (batch size) must be below 2^31 on the device path";
        return FFR_E_INVALID;
    }
    if (ctx->jit_ready && (ctx->jit_cfg.async || ctx->jit_cfg.affine) && chain_count > (1ULL << 30))
    {
        /* K1d hands out chains through a 32-bit counter: split very long launches */
        const u64 half = 1ULL << 30;
        int rc = launch_render(ctx,ds,chain_first,half,chain_len,0,base_seed,bv_limit);
        if (rc != FFR_OK)
            return rc;
        return launch_render(ctx,ds,chain_first + half,chain_count - half,chain_len,last_len,base_seed,bv_limit);
    }
    const u64 group_chains = ctx->jit_ready ? (ctx->jit_cfg.affine ? 32u : (u64)ctx->jit_cfg.ns) : (u64)FFR_TPB;
    const u64 groups = (chain_count + group_chains - 1) / group_chains;
    if (groups > 0xfffffff0ULL)
    {
        ctx->err = "too many chains in one launch";
        return FFR_E_INVALID;
    }
    u64 grid = (u64)ds.sm_count * (ctx->jit_ready ? ds.jit_blocks_per_sm : ds.blocks_per_sm);
    if (grid > groups)
        grid = groups;
    CK(cudaMemsetAsync(ds.d_counter,0,sizeof(unsigned int),ds.stream));
    if (ctx->jit_ready)
    {
        prm.rsl_scratch = ds.d_rsl_jit;
        void *args[] = {&prm};
        jit::Api &a = jit::api(true);
        /* K1d/K1e: scatter diagnostics (warp aggregation, discard, trace) live in a variant of the
           kernel that is compiled and loaded the first time one of them is asked for */
        CUfunction fn = ds.jmod.fn;
        if (ctx->scatter_mode != FFR_SCATTER_GLOBAL && (ctx->jit_cfg.async || ctx->jit_cfg.affine))
        {
            const int mrc = jit_load_modes(ctx,ds);
            if (mrc != FFR_OK)
                return mrc;
            fn = ds.jmod_modes.fn;
        }
        const CUresult r = a.LaunchKernel(fn,(unsigned)grid,1,1,(unsigned)ctx->jit_cfg.tpb,1,1,
            (unsigned)ctx->jit_smem,(CUstream)ds.stream,args,nullptr);
        if (r != CUDA_SUCCESS)
        {
            ctx->err = "cuLaunchKernel(ffr_jit_render): " + jit::cu_err(a,r);
            return FFR_E_CUDA;
        }
        if (fn == ds.jmod.fn)
        {
            const int frc = fold_tiles(ctx,ds);
            if (frc != FFR_OK)
                return frc;
        }
    }
    else
        ctx->kernel<<<(unsigned)grid,FFR_TPB,ctx->smem_bytes,ds.stream>>>(prm);
    CK(cudaGetLastError());
    ++ctx->launches;
    ds.dirty = true;
    return FFR_OK;
}

/* reset the per-call bad value list (render() clears it, buffer_renderer.hpp:277-278) */
int reset_bad(ffr_ctx *ctx, DeviceState &ds)
{
    CK(cudaSetDevice(ds.dev));
    CK(cudaMemsetAsync(&ds.d_stats->n_bad,0,sizeof(u64) + 2*sizeof(uint32_t),ds.stream));
    return FFR_OK;
}

int collect_stats(ffr_ctx *ctx, ffr_stats *out)
{
    memset(out,0,sizeof(*out));
    for (int i = 0; i < 3; ++i)
    {
        out->pt_min[i] = INFINITY;
        out->pt_max[i] = -INFINITY;
    }
    std::vector<DevStats> hs(1);
    for (DeviceState &ds : ctx->devs)
    {
        CK(cudaSetDevice(ds.dev));
        CK(cudaMemcpyAsync(&hs[0],ds.d_stats,sizeof(DevStats),cudaMemcpyDeviceToHost,ds.stream));
        CK(cudaStreamSynchronize(ds.stream));
        const DevStats &s = hs[0];
        if (s.abort & 0x100u)
        {
            ctx->err = "render kernel: slot queue watchdog tripped (internal error)";
            return FFR_E_CUDA;
        }
        out->s_iter += s.s_iter;
        out->s_plot += s.s_plot;
        for (uint32_t i = 0; i < ctx->num_xforms; ++i)
            out->xf_dist[ctx->json_ids[i]] += s.xf_dist[i];
        for (uint32_t i = 0; i < ctx->dims; ++i)
        {
            out->pt_min[i] = std::min(out->pt_min[i],ordered_to_f64(s.pt_min[i]));
            out->pt_max[i] = std::max(out->pt_max[i],ordered_to_f64(s.pt_max[i]));
        }
        const u64 nb = std::min<u64>(s.n_bad,FFR_MAX_BAD_RECORDED);
        for (u64 k = 0; k < nb; ++k)
        {
            if (out->n_bad + k < FFR_MAX_BAD_RECORDED)
            {
                out->bad_xf[out->n_bad + k] = s.bad_xf[k];
                for (int d = 0; d < 3; ++d)
                    out->bad_pt[out->n_bad + k][d] = s.bad_pt[k][d];
            }
        }
        out->n_bad += s.n_bad;
    }
    return FFR_OK;
}

int sync_all(ffr_ctx *ctx)
{
    for (DeviceState &ds : ctx->devs)
    {
        CK(cudaSetDevice(ds.dev));
        CK(cudaStreamSynchronize(ds.stream));
    }
    return FFR_OK;
}

bool aborted(ffr_ctx *ctx, u64 bv_limit, const ffr_stats &st)
{
    (void)ctx;
    return st.n_bad > bv_limit;
}

} // namespace

extern "C"
{

const char *ffr_cuda_version(void)
{
    return FFR_VERSION_STRING;
}

int ffr_cuda_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
    {
        cudaGetLastError();
        return 0;
    }
    return n;
}

uint64_t ffr_chain_seed(uint64_t base_seed, uint64_t chain_index)
{
    return splitmix64(base_seed + chain_index);
}

ffr_ctx *ffr_cuda_create(const ffr_flame_desc *desc, const int *devices, int ndev, char *err,
        size_t errlen)
{
    return ffr_cuda_create_ex(desc,devices,ndev,nullptr,err,errlen);
}

ffr_ctx *ffr_cuda_create_ex(const ffr_flame_desc *desc, const int *devices, int ndev,
        const ffr_options *opt, char *err, size_t errlen)
{
    /* FFR_TIMING=1: where the time to a ready context goes (stderr) */
    const bool timing = getenv("FFR_TIMING") && *getenv("FFR_TIMING") == '1';
    timespec t_begin;
    clock_gettime(CLOCK_MONOTONIC,&t_begin);
    auto lap = [&](const char *what)
    {
        if (!timing)
            return;
        timespec t;
        clock_gettime(CLOCK_MONOTONIC,&t);
        fprintf(stderr,"timing[%d]: create: %s at %.3f s\n",(int)getpid(),what,
            (double)(t.tv_sec - t_begin.tv_sec) + 1e-9*(double)(t.tv_nsec - t_begin.tv_nsec));
    };
    ffr_ctx *ctx = new ffr_ctx;
    std::string msg;
    auto fail = [&](const std::string& m) -> ffr_ctx*
    {
        if (err && errlen)
            snprintf(err,errlen,"%s",m.c_str());
        ffr_cuda_destroy(ctx);
        return nullptr;
    };
    memset(&ctx->opt,0,sizeof(ctx->opt));
    if (opt)
        memcpy(&ctx->opt,opt,std::min<size_t>(opt->struct_size ? opt->struct_size : sizeof(ffr_options),
            sizeof(ffr_options)));
    const bool f32 = desc && desc->elem_size == 4;
    if (!(f32 ? pack_blob<float>(ctx,desc,msg) : pack_blob<double>(ctx,desc,msg)))
        return fail(msg);
    if (ndev < 1)
        return fail("ffr_cuda_create(): need at least one device");
    if ((ctx->opt.external_buffer || ctx->opt.stream) && ndev != 1)
        return fail("ffr_cuda_create(): external buffer / stream need a single device context");
    int avail = ffr_cuda_device_count();
    lap("driver initialised (cudaGetDeviceCount)");
    if (avail < 1)
        return fail("ffr_cuda_create(): no CUDA device available; libffr_cuda has no CPU fallback");
    ctx->scatter_mode = ctx->opt.scatter_mode;
    if (ctx->scatter_mode == FFR_SCATTER_SMEM_TILE)
        return fail("ffr_cuda_create(): FFR_SCATTER_SMEM_TILE is not implemented (the shared memory of an SM "
                    "holds the chains' ISAAC state; see DESIGN.md, scatter strategies)");
    if (ctx->scatter_mode > FFR_SCATTER_TRACE)
        return fail("ffr_cuda_create(): unknown scatter mode");
    if (ctx->scatter_mode == FFR_SCATTER_AUTO)
        ctx->scatter_mode = FFR_SCATTER_GLOBAL;
    ctx->kernel = f32 ? pick_kernel<float>(ctx->dims,ctx->r,ctx->affine_only)
                      : pick_kernel<double>(ctx->dims,ctx->r,ctx->affine_only);
    ctx->smem_bytes = FFR_SMEM_RNG_BYTES_W(ctx->elem) + ctx->blob.size();
    /* regroup (K1b) when lanes would otherwise diverge over different op lists */
    ctx->regroup = false;
    if (ctx->opt.regroup != 1 && ctx->r <= 4 && ctx->num_xforms <= 31 &&
        ((ctx->opt.regroup == 2 && ctx->num_xforms >= 1) ||
         (ctx->distinct_oplists >= 2 && ctx->divergence_ratio > 2.2)))
    {
        size_t extra = 0;
        render_fn k = f32 ? pick_regroup<float>(ctx->dims,ctx->r,&extra)
                          : pick_regroup<double>(ctx->dims,ctx->r,&extra);
        if (k)
        {
            ctx->kernel = k;
            ctx->smem_bytes = extra + ctx->blob.size();
            ctx->regroup = true;
        }
    }
    ctx->devs.resize(ndev);
    if (ndev > 1)
    {
        /* creating a device's primary context takes a good part of a second: all of them at
           once, from one thread each (errors surface in the sequential set-up below) */
        std::vector<std::thread> warm;
        for (int i = 0; i < ndev; ++i)
        {
            const int dev = devices ? devices[i] : i;
            if (dev >= 0 && dev < avail)
                warm.emplace_back([dev]() { if (cudaSetDevice(dev) == cudaSuccess) cudaFree(0); });
        }
        for (std::thread &t : warm)
            t.join();
        cudaGetLastError();
        lap("primary contexts created");
    }
    for (int i = 0; i < ndev; ++i)
    {
        int dev = devices ? devices[i] : i;
        if (dev < 0 || dev >= avail)
            return fail("ffr_cuda_create(): device index out of range");
        int rc = setup_device(ctx,ctx->devs[i],dev,ctx->opt);
        if (rc != FFR_OK)
            return fail(ctx->err);
        lap("device set up (context, module, buffer)");
    }
    /* K1c/K1d (run-time compiled). opt.jit: 0 auto = compiled lazily by the first render call of
       >= FFR_JIT_MIN_SAMPLES samples, for flames with variations other than linear; 1 never;
       2 now, for any flame the kernel supports. FFR_JIT=0/1 in the environment overrides auto. */
    ctx->jit_mode = ctx->opt.jit;
    if (ctx->jit_mode == 0)
    {
        const char *e = getenv("FFR_JIT");
        if (e && *e == '0') ctx->jit_mode = 1;
        else if (e && *e == '1') ctx->jit_mode = 2;
    }
    /* pure-affine flames: K1e where it applies (r = 0, no final xform, ...); in auto mode the
       queue-scheduled kernel is not tried for them, K1 is the better kernel there */
    ctx->jit_eligible = ctx->jit_mode != 1 && jit_supported(ctx) &&
        (ctx->jit_mode == 2 || !ctx->affine_only || (ctx->r == 0 && !ctx->has_final));
    if (ctx->jit_mode == 2 && ctx->opt.jit == 2)
    {
        if (!ctx->jit_eligible)
            return fail("ffr_cuda_create(): the flame-specialised kernel supports <= 8 xforms and <= 4 colour dimensions");
        if (jit_activate(ctx) != FFR_OK)
            return fail("ffr_cuda_create(): run-time compilation failed: " + (ctx->jit_err.empty() ? ctx->err : ctx->jit_err));
        lap("flame-specialised kernel compiled/loaded");
    }
    else if (ctx->jit_mode == 2 && ctx->jit_eligible)
        jit_activate(ctx);   /* FFR_JIT=1: best effort, the interpreter kernels remain */
    return ctx;
}

void ffr_cuda_destroy(ffr_ctx *ctx)
{
    if (!ctx)
        return;
    for (DeviceState &ds : ctx->devs)
    {
        if (ds.dev < 0)
            continue;
        cudaSetDevice(ds.dev);
        if (ds.stream)
            cudaStreamSynchronize(ds.stream);
        if (ds.own_buffer && ds.buffer) cudaFree(ds.buffer);
        if (ds.d_blob) cudaFree(ds.d_blob);
        if (ds.d_colors) cudaFree(ds.d_colors);
        if (ds.d_stats) cudaFree(ds.d_stats);
        if (ds.d_counter) cudaFree(ds.d_counter);
        if (ds.d_scratch) cudaFree(ds.d_scratch);
        if (ds.d_rsl) cudaFree(ds.d_rsl);
        if (ds.d_rsl_jit) cudaFree(ds.d_rsl_jit);
        if (ds.d_acc) cudaFree(ds.d_acc);
        if (ds.d_dir) cudaFree(ds.d_dir);
        if (ds.jmod.mod) jit::unload(ds.jmod);
        if (ds.jmod_modes.mod) jit::unload(ds.jmod_modes);
        if (ds.d_stage) cudaFree(ds.d_stage);
        if (ds.copy_stream) cudaStreamDestroy(ds.copy_stream);
        if (ds.ev_copied) cudaEventDestroy(ds.ev_copied);
        if (ds.ev_added) cudaEventDestroy(ds.ev_added);
        if (ds.own_stream && ds.stream) cudaStreamDestroy(ds.stream);
    }
    delete ctx;
}

const char *ffr_cuda_last_error(const ffr_ctx *ctx)
{
    return ctx ? ctx->err.c_str() : g_create_err.c_str();
}

size_t ffr_cuda_buffer_bytes(const ffr_ctx *ctx)
{
    return ctx ? ctx->bytes : 0;
}

uint64_t ffr_cuda_buffer_cells(const ffr_ctx *ctx)
{
    return ctx ? ctx->cells : 0;
}

void *ffr_cuda_device_buffer(ffr_ctx *ctx, int dev_index)
{
    if (!ctx || dev_index < 0 || dev_index >= (int)ctx->devs.size())
        return nullptr;
    return ctx->devs[dev_index].buffer;
}

/* is this host pointer page-locked (cudaHostAlloc / cudaHostRegister)? */
static bool is_pinned(const void *host)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a,host) != cudaSuccess)
    {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

static int launch_add(ffr_ctx *ctx, DeviceState &ds, void *dst, const void *src, size_t n, size_t first)
{
    const unsigned grid = (unsigned)std::min<size_t>((n + 255)/256,(size_t)ds.sm_count*16);
    if (ctx->elem == 8)
        add_buffer_kernel<double><<<grid,256,0,ds.stream>>>((u64*)dst,(const u64*)src,n,ctx->cellsz,first);
    else
        add_buffer_kernel<float><<<grid,256,0,ds.stream>>>((unsigned int*)dst,(const unsigned int*)src,n,ctx->cellsz,first);
    ++ctx->launches;
    CK(cudaGetLastError());
    return FFR_OK;
}

static int ensure_stage(ffr_ctx *ctx, DeviceState &ds, size_t elems)
{
    if (ds.stage_elems >= elems)
        return FFR_OK;
    CK(cudaStreamSynchronize(ds.stream));
    if (ds.d_stage)
        cudaFree(ds.d_stage);
    ds.d_stage = nullptr;
    ds.stage_elems = 0;
    ds.stage_busy = false;
    CK(cudaMalloc(&ds.d_stage,elems*ctx->elem));
    ds.stage_elems = elems;
    return FFR_OK;
}

int ffr_cuda_add_buffer(ffr_ctx *ctx, const void *host, size_t bytes)
{
    if (!ctx || !host)
        return FFR_E_INVALID;
    if (bytes != ctx->bytes)
    {
        ctx->err = "BufferRenderer::addBuffer(): sizes do not match";
        return FFR_E_INVALID;
    }
    DeviceState &ds = ctx->devs[0];
    CK(cudaSetDevice(ds.dev));
    /* stage in chunks of 128 MiB so that a 1 GiB -i file does not double the footprint; the staging
       buffer is kept for the next call. Stream order protects its reuse (the copy of chunk k+1
       starts after the add of chunk k); ONE synchronisation at the end, because the host buffer is
       only borrowed. */
    const size_t eb = ctx->elem;
    const size_t n_elems = bytes/eb;
    const size_t chunk_elems = ((size_t)1 << 27) / eb;
    const size_t chunk = std::min(n_elems,chunk_elems - (chunk_elems % ctx->cellsz));
    int rc = ensure_stage(ctx,ds,chunk);
    if (rc != FFR_OK)
        return rc;
    if (ds.stage_busy)
    {
        CK(cudaStreamWaitEvent(ds.stream,ds.ev_added,0));
        ds.stage_busy = false;
    }
    for (size_t off = 0; off < n_elems; off += chunk)
    {
        const size_t n = std::min(chunk,n_elems - off);
        CK(cudaMemcpyAsync(ds.d_stage,(const char*)host + off*eb,n*eb,cudaMemcpyHostToDevice,ds.stream));
        rc = launch_add(ctx,ds,(char*)ds.buffer + off*eb,ds.d_stage,n,off);
        if (rc != FFR_OK)
            return rc;
    }
    CK(cudaStreamSynchronize(ds.stream));
    return FFR_OK;
}

static int single_device(ffr_ctx *ctx, const char *what)
{
    if (ctx->devs.size() != 1)
    {
        ctx->err = std::string(what) + " needs a single device context";
        return FFR_E_INVALID;
    }
    return FFR_OK;
}

int ffr_cuda_add_buffer_async(ffr_ctx *ctx, const void *pinned_host, size_t bytes)
{
    if (!ctx || !pinned_host)
        return FFR_E_INVALID;
    if (single_device(ctx,"add_buffer_async") != FFR_OK)
        return FFR_E_INVALID;
    if (bytes != ctx->bytes)
    {
        ctx->err = "BufferRenderer::addBuffer(): sizes do not match";
        return FFR_E_INVALID;
    }
    DeviceState &ds = ctx->devs[0];
    CK(cudaSetDevice(ds.dev));
    if (!is_pinned(pinned_host))
    {
        ctx->err = "add_buffer_async: the host buffer must be page-locked";
        return FFR_E_INVALID;
    }
    /* the whole buffer is uploaded by a copy engine on a stream of its own -- it overlaps whatever
       kernels run meanwhile, a render of another context in particular -- into a full-size staging
       buffer (this interface trades memory for overlap); the add kernel then runs on the
       context's stream, ordered after the copy by an event, and the next upload after this add */
    int rc = ensure_stage(ctx,ds,bytes/ctx->elem);
    if (rc != FFR_OK)
        return rc;
    if (!ds.copy_stream)
    {
        CK(cudaStreamCreateWithFlags(&ds.copy_stream,cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&ds.ev_copied,cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ds.ev_added,cudaEventDisableTiming));
    }
    if (ds.stage_busy)
        CK(cudaStreamWaitEvent(ds.copy_stream,ds.ev_added,0));
    CK(cudaMemcpyAsync(ds.d_stage,pinned_host,bytes,cudaMemcpyHostToDevice,ds.copy_stream));
    CK(cudaEventRecord(ds.ev_copied,ds.copy_stream));
    CK(cudaStreamWaitEvent(ds.stream,ds.ev_copied,0));
    rc = launch_add(ctx,ds,ds.buffer,ds.d_stage,bytes/ctx->elem,0);
    if (rc != FFR_OK)
        return rc;
    CK(cudaEventRecord(ds.ev_added,ds.stream));
    ds.stage_busy = true;
    return FFR_OK;
}

int ffr_cuda_read_buffer_async(ffr_ctx *ctx, void *pinned_host, size_t bytes)
{
    if (!ctx || !pinned_host)
        return FFR_E_INVALID;
    if (single_device(ctx,"read_buffer_async") != FFR_OK)
        return FFR_E_INVALID;
    if (bytes != ctx->bytes)
    {
        ctx->err = "read_buffer: size mismatch";
        return FFR_E_INVALID;
    }
    DeviceState &ds = ctx->devs[0];
    CK(cudaSetDevice(ds.dev));
    if (!is_pinned(pinned_host))
    {
        ctx->err = "read_buffer_async: the host buffer must be page-locked";
        return FFR_E_INVALID;
    }
    CK(cudaMemcpyAsync(pinned_host,ds.buffer,bytes,cudaMemcpyDeviceToHost,ds.stream));
    return FFR_OK;
}

int ffr_cuda_clear_buffer_async(ffr_ctx *ctx)
{
    if (!ctx)
        return FFR_E_INVALID;
    if (single_device(ctx,"clear_buffer_async") != FFR_OK)
        return FFR_E_INVALID;
    DeviceState &ds = ctx->devs[0];
    CK(cudaSetDevice(ds.dev));
    CK(cudaMemsetAsync(ds.buffer,0,ctx->bytes,ds.stream));
    ds.dirty = false;
    return FFR_OK;
}

int ffr_cuda_clear_buffer(ffr_ctx *ctx)
{
    if (!ctx)
        return FFR_E_INVALID;
    for (DeviceState &ds : ctx->devs)
    {
        CK(cudaSetDevice(ds.dev));
        CK(cudaMemsetAsync(ds.buffer,0,ctx->bytes,ds.stream));
        ds.dirty = false;
    }
    return sync_all(ctx);
}

int ffr_cuda_render_chains_async(ffr_ctx *ctx, uint64_t chain_first, uint64_t chain_count,
        uint64_t chain_len, uint64_t last_len, uint64_t base_seed, uint64_t bv_limit)
{
    if (!ctx)
        return FFR_E_INVALID;
    if (ctx->devs.size() != 1)
    {
        ctx->err = "render_chains_async needs a single device context";
        return FFR_E_INVALID;
    }
    if (chain_len == 0 || last_len > chain_len)
    {
        ctx->err = "BufferRenderer::render(): batch size must be positive";
        return FFR_E_INVALID;
    }
    return launch_render(ctx,ctx->devs[0],chain_first,chain_count,chain_len,last_len,base_seed,bv_limit);
}

int ffr_cuda_sync(ffr_ctx *ctx)
{
    if (!ctx)
        return FFR_E_INVALID;
    return sync_all(ctx);
}

int ffr_cuda_get_stats(ffr_ctx *ctx, ffr_stats *stats)
{
    if (!ctx || !stats)
        return FFR_E_INVALID;
    return collect_stats(ctx,stats);
}

uint64_t ffr_cuda_resident_chains(const ffr_ctx *ctx)
{
    if (!ctx || ctx->devs.empty())
        return 0;
    if (ctx->jit_ready)
        return (uint64_t)ctx->devs[0].sm_count * ctx->devs[0].jit_blocks_per_sm * ctx->jit_cfg.ns;
    return (uint64_t)ctx->devs[0].sm_count * ctx->devs[0].blocks_per_sm * FFR_TPB;
}

uint64_t ffr_cuda_launch_count(const ffr_ctx *ctx)
{
    return ctx ? ctx->launches : 0;
}

int ffr_cuda_render_chains(ffr_ctx *ctx, uint64_t chain_first, uint64_t chain_count,
        uint64_t chain_len, uint64_t last_len, uint64_t base_seed, uint64_t bv_limit,
        ffr_stats *stats)
{
    if (!ctx)
        return FFR_E_INVALID;
    if (chain_len == 0 || last_len > chain_len)
    {
        ctx->err = "BufferRenderer::render(): batch size must be positive";
        return FFR_E_INVALID;
    }
    const size_t nd = ctx->devs.size();
    jit_maybe(ctx,chain_count*chain_len);
    for (DeviceState &ds : ctx->devs)
    {
        int rc = reset_bad(ctx,ds);
        if (rc != FFR_OK)
            return rc;
    }
    /* contiguous chain ranges, one per device */
    u64 per = (chain_count + nd - 1) / nd;
    for (size_t i = 0; i < nd; ++i)
    {
        u64 first = std::min<u64>(per*i,chain_count);
        u64 count = std::min<u64>(per,chain_count - first);
        bool has_last = (first + count == chain_count);
        int rc = launch_render(ctx,ctx->devs[i],chain_first + first,count,chain_len,
            has_last ? last_len : 0,base_seed,bv_limit);
        if (rc != FFR_OK)
            return rc;
    }
    int rc = sync_all(ctx);
    if (rc != FFR_OK)
        return rc;
    ffr_stats local;
    ffr_stats *st = stats ? stats : &local;
    rc = collect_stats(ctx,st);
    if (rc != FFR_OK)
        return rc;
    return aborted(ctx,bv_limit,*st) ? FFR_BAD_VALUES : FFR_OK;
}

int ffr_cuda_render(ffr_ctx *ctx, uint64_t samples, uint64_t chain_len, uint64_t base_seed,
        uint64_t bv_limit, ffr_progress_cb cb, void *user, ffr_stats *stats)
{
    if (!ctx)
        return FFR_E_INVALID;
    for (DeviceState &ds : ctx->devs)
    {
        int rc = reset_bad(ctx,ds);
        if (rc != FFR_OK)
            return rc;
    }
    if (samples == 0) /* buffer_renderer.hpp:279-280 */
    {
        if (stats)
            return collect_stats(ctx,stats);
        return FFR_OK;
    }
    if (chain_len < 256) /* :283-285 */
    {
        ctx->err = "BufferRenderer::render(): batch size too small";
        return FFR_E_INVALID;
    }
    jit_maybe(ctx,samples);
    const u64 chains = (samples + chain_len - 1) / chain_len;
    const u64 last = samples - (chains - 1)*chain_len;
    const u64 last_len = (last == chain_len) ? 0 : last;
    if (!cb)
        return ffr_cuda_render_chains(ctx,0,chains,chain_len,last_len,base_seed,bv_limit,stats);
    /* with a progress callback: a few launches, callback after each from this thread. Every
       launch gives every device at least four waves of its resident chains (jit_maybe above has
       settled which kernel runs): with fewer chains than chain slots the device idles -- cfg5 of
       the baseline (1e11 samples, 8 devices, 8192-sample chains) ran at 7.9e10 samples/s with
       32 segments of 47 684 chains per device against 151 552 slots, 1.85e11 with segments sized
       like this. */
    const u64 wave = std::max<u64>(1,ffr_cuda_resident_chains(ctx))*(u64)ctx->devs.size();
    const u64 segs = std::min<u64>(32,std::max<u64>(1,chains / (wave*4)));
    const u64 per = (chains + segs - 1) / segs;
    int rc = FFR_OK;
    ffr_stats local;
    ffr_stats *st = stats ? stats : &local;
    for (u64 first = 0; first < chains; first += per)
    {
        u64 count = std::min<u64>(per,chains - first);
        bool has_last = (first + count == chains);
        const size_t nd = ctx->devs.size();
        u64 dper = (count + nd - 1) / nd;
        for (size_t i = 0; i < nd; ++i)
        {
            u64 f = std::min<u64>(dper*i,count);
            u64 c = std::min<u64>(dper,count - f);
            rc = launch_render(ctx,ctx->devs[i],first + f,c,chain_len,
                (has_last && f + c == count) ? last_len : 0,base_seed,bv_limit);
            if (rc != FFR_OK)
                return rc;
        }
        rc = sync_all(ctx);
        if (rc != FFR_OK)
            return rc;
        rc = collect_stats(ctx,st);
        if (rc != FFR_OK)
            return rc;
        cb(user,first + count,chains);
        if (aborted(ctx,bv_limit,*st))
            return FFR_BAD_VALUES;
    }
    return FFR_OK;
}

/* launch K2d on `ds` for elements [first, first + n) of the buffer */
static int launch_reduce_slices(ffr_ctx *ctx, DeviceState &ds, void *dst, const PeerSlices &ps, int n_src,
        u64 first, u64 n)
{
    if (n == 0 || n_src == 0)
        return FFR_OK;
    const unsigned grid = (unsigned)std::min<u64>((n + 255)/256,(u64)ds.sm_count*8);
    if (ctx->elem == 8)
        reduce_slices_kernel<double><<<grid,256,0,ds.stream>>>((u64*)dst,ps,n_src,first,n,ctx->cellsz);
    else
        reduce_slices_kernel<float><<<grid,256,0,ds.stream>>>((unsigned int*)dst,ps,n_src,first,n,ctx->cellsz);
    ++ctx->launches;
    CK(cudaGetLastError());
    return FFR_OK;
}

int ffr_cuda_reduce(ffr_ctx *ctx)
{
    if (!ctx)
        return FFR_E_INVALID;
    const size_t nd = ctx->devs.size();
    if (nd < 2)
        return FFR_OK;
    bool any = false;
    for (size_t i = 1; i < nd; ++i)
        any |= ctx->devs[i].dirty;
    if (!any)
        return FFR_OK;
    if (nd > FFR_MAX_PEERS + 1)
    {
        ctx->err = "ffr_cuda_reduce(): at most 16 devices";
        return FFR_E_INVALID;
    }
    int rc = sync_all(ctx);
    if (rc != FFR_OK)
        return rc;
    /* peer access between every pair, decided once per context */
    if (!ctx->peer_probed)
    {
        ctx->peer_probed = true;
        ctx->peer_enabled = true;
        for (size_t i = 0; i < nd && ctx->peer_enabled; ++i)
            for (size_t j = 0; j < nd; ++j)
            {
                if (i == j)
                    continue;
                int can = 0;
                CK(cudaDeviceCanAccessPeer(&can,ctx->devs[i].dev,ctx->devs[j].dev));
                if (!can)
                {
                    ctx->peer_enabled = false;
                    break;
                }
            }
        if (ctx->peer_enabled)
            for (size_t i = 0; i < nd; ++i)
            {
                CK(cudaSetDevice(ctx->devs[i].dev));
                for (size_t j = 0; j < nd; ++j)
                {
                    if (i == j)
                        continue;
                    const cudaError_t e = cudaDeviceEnablePeerAccess(ctx->devs[j].dev,0);
                    if (e == cudaErrorPeerAccessAlreadyEnabled)
                        cudaGetLastError();
                    else
                        CK(e);
                }
            }
    }
    DeviceState &d0 = ctx->devs[0];
    const u64 n_elems = ctx->bytes/ctx->elem;
    if (!ctx->peer_enabled)
    {
        /* no P2P path between some pair: device 0 adds one staged copy after the other (the
           staging buffer is the one add_buffer keeps) */
        CK(cudaSetDevice(d0.dev));
        const u64 chunk = std::min<u64>(n_elems,((u64)1 << 27)/ctx->elem);
        if (d0.stage_elems < chunk)
        {
            if (d0.d_stage)
                cudaFree(d0.d_stage);
            d0.d_stage = nullptr;
            d0.stage_elems = 0;
            CK(cudaMalloc(&d0.d_stage,chunk*ctx->elem));
            d0.stage_elems = chunk;
        }
        for (size_t i = 1; i < nd; ++i)
        {
            if (!ctx->devs[i].dirty)
                continue;
            for (u64 off = 0; off < n_elems; off += chunk)
            {
                const u64 n = std::min<u64>(chunk,n_elems - off);
                CK(cudaMemcpyPeerAsync(d0.d_stage,d0.dev,(const char*)ctx->devs[i].buffer + off*ctx->elem,
                    ctx->devs[i].dev,n*ctx->elem,d0.stream));
                PeerSlices ps;
                memset(&ps,0,sizeof(ps));
                ps.src[0] = d0.d_stage;
                rc = launch_reduce_slices(ctx,d0,(char*)d0.buffer + off*ctx->elem,ps,1,off,n);
                if (rc != FFR_OK)
                    return rc;
            }
        }
        CK(cudaStreamSynchronize(d0.stream));
    }
    else
    {
        /* reduce-scatter: device d sums slice d of every private buffer into its own, all devices
           at once; then device 0 gathers the finished slices. Slices are whole cells. */
        const u64 cells_per = (ctx->cells + nd - 1)/nd;
        std::vector<cudaEvent_t> done(nd,nullptr);
        struct EvGuard
        {
            std::vector<cudaEvent_t> &v;
            ~EvGuard() { for (cudaEvent_t e : v) if (e) cudaEventDestroy(e); }
        } guard{done};
        for (size_t d = 0; d < nd; ++d)
        {
            DeviceState &ds = ctx->devs[d];
            CK(cudaSetDevice(ds.dev));
            const u64 c0 = std::min<u64>(cells_per*d,ctx->cells), c1 = std::min<u64>(c0 + cells_per,ctx->cells);
            const u64 first = c0*ctx->cellsz, n = (c1 - c0)*ctx->cellsz;
            PeerSlices ps;
            memset(&ps,0,sizeof(ps));
            int k = 0;
            for (size_t j = 0; j < nd; ++j)
                if (j != d && (ctx->devs[j].dirty || j == 0))
                    ps.src[k++] = (const char*)ctx->devs[j].buffer + first*ctx->elem;
            rc = launch_reduce_slices(ctx,ds,(char*)ds.buffer + first*ctx->elem,ps,k,first,n);
            if (rc != FFR_OK)
                return rc;
            CK(cudaEventCreateWithFlags(&done[d],cudaEventDisableTiming));
            CK(cudaEventRecord(done[d],ds.stream));
        }
        CK(cudaSetDevice(d0.dev));
        for (size_t d = 1; d < nd; ++d)
        {
            const u64 c0 = std::min<u64>(cells_per*d,ctx->cells), c1 = std::min<u64>(c0 + cells_per,ctx->cells);
            const u64 first = c0*ctx->cellsz, n = (c1 - c0)*ctx->cellsz;
            if (n == 0)
                continue;
            CK(cudaStreamWaitEvent(d0.stream,done[d],0));
            CK(cudaMemcpyPeerAsync((char*)d0.buffer + first*ctx->elem,d0.dev,
                (const char*)ctx->devs[d].buffer + first*ctx->elem,ctx->devs[d].dev,n*ctx->elem,d0.stream));
        }
        /* every device must be done reading device 0's slices before anything else touches them */
        for (size_t d = 1; d < nd; ++d)
            CK(cudaStreamWaitEvent(d0.stream,done[d],0));
        CK(cudaStreamSynchronize(d0.stream));
    }
    /* the peers' samples now live in device 0: clear them so a later reduce adds nothing twice */
    for (size_t i = 1; i < nd; ++i)
    {
        DeviceState &ds = ctx->devs[i];
        CK(cudaSetDevice(ds.dev));
        CK(cudaMemsetAsync(ds.buffer,0,ctx->bytes,ds.stream));
        ds.dirty = false;
    }
    return sync_all(ctx);
}

int ffr_cuda_sum_device_slices(ffr_ctx *ctx, void *dst, const void *const *srcs, int n_src,
        uint64_t first_elem, uint64_t n_elems)
{
    if (!ctx || !dst || !srcs || n_src < 0 || n_src > FFR_MAX_PEERS)
        return FFR_E_INVALID;
    if (ctx->devs.size() != 1)
    {
        ctx->err = "sum_device_slices needs a single device context";
        return FFR_E_INVALID;
    }
    DeviceState &ds = ctx->devs[0];
    CK(cudaSetDevice(ds.dev));
    PeerSlices ps;
    memset(&ps,0,sizeof(ps));
    for (int k = 0; k < n_src; ++k)
        ps.src[k] = srcs[k];
    return launch_reduce_slices(ctx,ds,dst,ps,n_src,first_elem,n_elems);
}

int ffr_cuda_ipc_export(ffr_ctx *ctx, void *handle)
{
    if (!ctx || !handle)
        return FFR_E_INVALID;
    if (single_device(ctx,"ipc_export") != FFR_OK)
        return FFR_E_INVALID;
    DeviceState &ds = ctx->devs[0];
    if (!ds.own_buffer)
    {
        ctx->err = "ipc_export: the buffer belongs to the caller (external_buffer)";
        return FFR_E_INVALID;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == FFR_IPC_HANDLE_BYTES,"CUDA IPC handle size");
    CK(cudaSetDevice(ds.dev));
    CK(cudaStreamSynchronize(ds.stream));
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h,ds.buffer));
    memcpy(handle,&h,sizeof(h));
    return FFR_OK;
}

int ffr_cuda_ipc_add(ffr_ctx *ctx, const void *handles, int n)
{
    if (!ctx || (!handles && n > 0) || n < 0 || n > FFR_MAX_PEERS)
        return FFR_E_INVALID;
    if (single_device(ctx,"ipc_add") != FFR_OK)
        return FFR_E_INVALID;
    if (n == 0)
        return FFR_OK;
    DeviceState &ds = ctx->devs[0];
    CK(cudaSetDevice(ds.dev));
    PeerSlices ps;
    memset(&ps,0,sizeof(ps));
    void *opened[FFR_MAX_PEERS] = {nullptr};
    int rc = FFR_OK;
    for (int k = 0; k < n && rc == FFR_OK; ++k)
    {
        cudaIpcMemHandle_t h;
        memcpy(&h,(const char*)handles + (size_t)k*FFR_IPC_HANDLE_BYTES,sizeof(h));
        if (!cuda_ok(ctx,cudaIpcOpenMemHandle(&opened[k],h,cudaIpcMemLazyEnablePeerAccess),"cudaIpcOpenMemHandle"))
            rc = FFR_E_CUDA;
        ps.src[k] = opened[k];
    }
    if (rc == FFR_OK)
        rc = launch_reduce_slices(ctx,ds,ds.buffer,ps,n,0,ctx->bytes/ctx->elem);
    if (rc == FFR_OK && !cuda_ok(ctx,cudaStreamSynchronize(ds.stream),"ipc_add sync"))
        rc = FFR_E_CUDA;
    for (int k = 0; k < n; ++k)
        if (opened[k])
            cudaIpcCloseMemHandle(opened[k]);
    return rc;
}

int ffr_cuda_read_buffer(ffr_ctx *ctx, void *host, size_t bytes)
{
    if (!ctx || !host)
        return FFR_E_INVALID;
    if (bytes != ctx->bytes)
    {
        ctx->err = "read_buffer: size mismatch";
        return FFR_E_INVALID;
    }
    int rc = ffr_cuda_reduce(ctx);
    if (rc != FFR_OK)
        return rc;
    DeviceState &ds = ctx->devs[0];
    CK(cudaSetDevice(ds.dev));
    CK(cudaMemcpyAsync(host,ds.buffer,bytes,cudaMemcpyDeviceToHost,ds.stream));
    CK(cudaStreamSynchronize(ds.stream));
    return FFR_OK;
}

int ffr_cuda_histogram_sum_max(ffr_ctx *ctx, uint64_t *sum, uint64_t *max)
{
    if (!ctx)
        return FFR_E_INVALID;
    int rc = ffr_cuda_reduce(ctx);
    if (rc != FFR_OK)
        return rc;
    DeviceState &ds = ctx->devs[0];
    CK(cudaSetDevice(ds.dev));
    CK(cudaMemsetAsync(ds.d_scratch,0,2*sizeof(u64),ds.stream));
    unsigned grid = (unsigned)std::min<u64>((ctx->cells + 255)/256,(u64)ds.sm_count*16);
    if (ctx->elem == 8)
        hist_sum_max_kernel<u64><<<grid,256,0,ds.stream>>>((const u64*)ds.buffer,ctx->cells,ctx->cellsz,
            ds.d_scratch,ds.d_scratch+1);
    else
        hist_sum_max_kernel<unsigned int><<<grid,256,0,ds.stream>>>((const unsigned int*)ds.buffer,
            ctx->cells,ctx->cellsz,ds.d_scratch,ds.d_scratch+1);
    ++ctx->launches;
    CK(cudaGetLastError());
    u64 h[2];
    CK(cudaMemcpyAsync(h,ds.d_scratch,sizeof(h),cudaMemcpyDeviceToHost,ds.stream));
    CK(cudaStreamSynchronize(ds.stream));
    if (sum) *sum = h[0];
    if (max) *max = h[1];
    return FFR_OK;
}

int ffr_cuda_tonemap(ffr_ctx *ctx, int mode, int bits, double gamma, void *pixels, size_t bytes,
        ffr_tonemap_info *info)
{
    if (!ctx || !pixels)
        return FFR_E_INVALID;
    if (ctx->dims != 2)
    {
        ctx->err = "only 2D flames supported"; /* ffr_img.cpp:123-127 */
        return FFR_E_INVALID;
    }
    if (!(gamma >= 1e-20)) /* ffr_img.cpp:88-89 */
    {
        ctx->err = "gamma too small";
        return FFR_E_INVALID;
    }
    if (bits != 8 && bits != 16) /* :94-95 */
    {
        ctx->err = "bits per channel must be 8 or 16";
        return FFR_E_INVALID;
    }
    if (mode != FFR_TONE_MONO && mode != FFR_TONE_GRAY && mode != FFR_TONE_RGB)
    {
        ctx->err = "no coloring flag";
        return FFR_E_INVALID;
    }
    if (mode == FFR_TONE_RGB && ctx->r != 3) /* :282-283 */
    {
        ctx->err = "buffer must use 3 color dimensions";
        return FFR_E_INVALID;
    }
    if (mode == FFR_TONE_MONO)
        bits = 8; /* renderGrayImage<u8>, :264 */
    const uint32_t channels = (mode == FFR_TONE_RGB) ? 3 : 1;
    const size_t need = (size_t)ctx->cells*channels*(bits/8);
    if (bytes != need)
    {
        ctx->err = "tonemap: pixel buffer size mismatch";
        return FFR_E_INVALID;
    }
    int rc = ffr_cuda_reduce(ctx);
    if (rc != FFR_OK)
        return rc;
    DeviceState &ds = ctx->devs[0];
    CK(cudaSetDevice(ds.dev));
    const u64 init[2] = {~0ULL,0ULL};
    CK(cudaMemcpyAsync(ds.d_scratch,init,sizeof(init),cudaMemcpyHostToDevice,ds.stream));
    unsigned grid = (unsigned)std::min<u64>((ctx->cells + 255)/256,(u64)ds.sm_count*16);
    if (ctx->elem == 8)
        hist_min_max_kernel<u64><<<grid,256,0,ds.stream>>>((const u64*)ds.buffer,ctx->cells,ctx->cellsz,
            ds.d_scratch,ds.d_scratch+1);
    else
        hist_min_max_kernel<unsigned int><<<grid,256,0,ds.stream>>>((const unsigned int*)ds.buffer,
            ctx->cells,ctx->cellsz,ds.d_scratch,ds.d_scratch+1);
    ++ctx->launches;
    CK(cudaGetLastError());
    u64 mm[2];
    CK(cudaMemcpyAsync(mm,ds.d_scratch,sizeof(mm),cudaMemcpyDeviceToHost,ds.stream));
    CK(cudaStreamSynchronize(ds.stream));
    if (info)
    {
        memset(info,0,sizeof(*info));
        info->hist_min = mm[0];
        info->hist_max = mm[1];
        info->scaler_min = log(1 + (double)mm[0]);
        info->scaler_max = log(1 + (double)mm[1]);
        info->width = ctx->size0;
        info->height = ctx->size1;
        info->channels = channels;
        info->bits = (uint32_t)bits;
    }
    if (log(1 + (double)mm[1]) < 1e-20) /* :231-232 */
    {
        ctx->err = "histogram is (probably) empty";
        return FFR_E_INVALID;
    }
    DevTemp pix;
    CK(pix.alloc(need));
    void *d_pix = pix.p;
    /* num_t gp = 1.0 / arg_gamma (:235); arg_gamma is a num_t */
    if (ctx->elem == 8)
    {
        const double gp = 1.0 / gamma;
        if (bits == 8)
            tonemap_kernel<double,unsigned char><<<grid,256,0,ds.stream>>>((const u64*)ds.buffer,ctx->cells,
                ctx->cellsz,mode,mm[1],gp,(unsigned char*)d_pix);
        else
            tonemap_kernel<double,unsigned short><<<grid,256,0,ds.stream>>>((const u64*)ds.buffer,ctx->cells,
                ctx->cellsz,mode,mm[1],gp,(unsigned short*)d_pix);
    }
    else
    {
        const float gp = (float)(1.0 / (float)gamma);
        if (bits == 8)
            tonemap_kernel<float,unsigned char><<<grid,256,0,ds.stream>>>((const unsigned int*)ds.buffer,
                ctx->cells,ctx->cellsz,mode,mm[1],gp,(unsigned char*)d_pix);
        else
            tonemap_kernel<float,unsigned short><<<grid,256,0,ds.stream>>>((const unsigned int*)ds.buffer,
                ctx->cells,ctx->cellsz,mode,mm[1],gp,(unsigned short*)d_pix);
    }
    ++ctx->launches;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(pixels,d_pix,need,cudaMemcpyDeviceToHost,ds.stream));
    CK(cudaStreamSynchronize(ds.stream));
    return FFR_OK;
}

int ffr_cuda_iterate_points(ffr_ctx *ctx, int64_t xf_index, uint64_t n, const uint64_t *seeds,
        const double *pts_in, double *pts_out)
{
    if (!ctx || !seeds || !pts_in || !pts_out)
        return FFR_E_INVALID;
    int slot;
    if (xf_index < 0)
    {
        if (!ctx->has_final)
        {
            ctx->err = "no final xform";
            return FFR_E_INVALID;
        }
        slot = (int)ctx->num_xforms;
    }
    else
    {
        if ((u64)xf_index >= ctx->num_xforms)
        {
            ctx->err = "xform index out of range";
            return FFR_E_INVALID;
        }
        slot = (int)xf_index;
    }
    if (n == 0)
        return FFR_OK;
    DeviceState &ds = ctx->devs[0];
    CK(cudaSetDevice(ds.dev));
    DevTemp t_seeds, t_in, t_out;
    const size_t pb = n*ctx->dims*sizeof(double);
    CK(t_seeds.alloc(n*8));
    CK(t_in.alloc(pb));
    CK(t_out.alloc(pb));
    u64 *d_seeds = t_seeds.as<u64>();
    double *d_in = t_in.as<double>(), *d_out = t_out.as<double>();
    CK(cudaMemcpyAsync(d_seeds,seeds,n*8,cudaMemcpyHostToDevice,ds.stream));
    CK(cudaMemcpyAsync(d_in,pts_in,pb,cudaMemcpyHostToDevice,ds.stream));
    const unsigned grid = (unsigned)((n + FFR_TPB - 1)/FFR_TPB);
    const uint32_t bb = (uint32_t)ctx->blob.size();
    const size_t sm = FFR_SMEM_RNG_BYTES_W(ctx->elem) + ctx->blob.size();
#define ITER_CASE(TT,DD) do { \
        CK(cudaFuncSetAttribute((const void*)iterate_points_kernel<TT,DD>, \
            cudaFuncAttributeMaxDynamicSharedMemorySize,(int)sm)); \
        iterate_points_kernel<TT,DD><<<grid,FFR_TPB,sm,ds.stream>>>(ds.d_blob,bb,slot,n,d_seeds,d_in,d_out); \
    } while (0)
    if (ctx->elem == 8)
    {
        if (ctx->dims == 1) ITER_CASE(double,1);
        else if (ctx->dims == 2) ITER_CASE(double,2);
        else ITER_CASE(double,3);
    }
    else
    {
        if (ctx->dims == 1) ITER_CASE(float,1);
        else if (ctx->dims == 2) ITER_CASE(float,2);
        else ITER_CASE(float,3);
    }
#undef ITER_CASE
    ++ctx->launches;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(pts_out,d_out,pb,cudaMemcpyDeviceToHost,ds.stream));
    CK(cudaStreamSynchronize(ds.stream));
    return FFR_OK;
}

int ffr_cuda_isaac_words(ffr_ctx *ctx, uint64_t seed, uint64_t n, uint64_t *out)
{
    if (!ctx || !out)
        return FFR_E_INVALID;
    if (n == 0)
        return FFR_OK;
    DeviceState &ds = ctx->devs[0];
    CK(cudaSetDevice(ds.dev));
    DevTemp t_out;
    CK(t_out.alloc(n*8));
    u64 *d_out = t_out.as<u64>();
    if (ctx->elem == 8)
    {
        CK(cudaFuncSetAttribute((const void*)isaac_words_kernel<double>,
            cudaFuncAttributeMaxDynamicSharedMemorySize,FFR_SMEM_RNG_BYTES_W(8)));
        isaac_words_kernel<double><<<1,FFR_TPB,FFR_SMEM_RNG_BYTES_W(8),ds.stream>>>(seed,n,d_out);
    }
    else
        isaac_words_kernel<float><<<1,FFR_TPB,FFR_SMEM_RNG_BYTES_W(4),ds.stream>>>(seed,n,d_out);
    ++ctx->launches;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out,d_out,n*8,cudaMemcpyDeviceToHost,ds.stream));
    CK(cudaStreamSynchronize(ds.stream));
    return FFR_OK;
}


int ffr_cuda_jit_info(const ffr_ctx *ctx, ffr_jit_info *info)
{
    if (!ctx || !info)
        return FFR_E_INVALID;
    memset(info,0,sizeof(*info));
    info->active = ctx->jit_ready ? 1u : 0u;
    info->eligible = ctx->jit_eligible ? 1u : 0u;
    info->failed = ctx->jit_failed ? 1u : 0u;
    info->from_cache = ctx->jit_cached ? 1u : 0u;
    info->threads_per_block = (uint32_t)ctx->jit_cfg.tpb;
    info->slots_per_block = (uint32_t)ctx->jit_cfg.ns;
    info->blocks_per_sm = ctx->devs.empty() ? 0u : (uint32_t)ctx->devs[0].jit_blocks_per_sm;
    info->registers = ctx->devs.empty() ? 0u : (uint32_t)ctx->devs[0].jmod.regs;
    info->smem_bytes = (uint64_t)ctx->jit_smem;
    info->cubin_bytes = (uint64_t)ctx->jit_cubin.size();
    info->source_bytes = (uint64_t)ctx->jit_source.size();
    info->compile_seconds = ctx->jit_compile_s;
    snprintf(info->message,sizeof(info->message),"%s",ctx->jit_err.empty() ? ctx->jit_note.c_str() : ctx->jit_err.c_str());
    return FFR_OK;
}

int ffr_cuda_jit_enable(ffr_ctx *ctx)
{
    if (!ctx)
        return FFR_E_INVALID;
    if (ctx->jit_mode == 1 || !jit_supported(ctx))
    {
        ctx->err = "the flame-specialised kernel is disabled or does not support this flame";
        return FFR_E_UNSUPPORTED;
    }
    ctx->jit_eligible = true;
    const int rc = jit_activate(ctx);
    if (rc != FFR_OK && ctx->err.empty())
        ctx->err = ctx->jit_err;
    return rc;
}

size_t ffr_cuda_jit_source(const ffr_ctx *ctx, char *buf, size_t buflen)
{
    if (!ctx)
        return 0;
    if (buf && buflen)
        snprintf(buf,buflen,"%s",ctx->jit_source.c_str());
    return ctx->jit_source.size();
}

int ffr_cuda_jit_compile(const ffr_flame_desc *desc, char *source, size_t source_len,
        size_t *cubin_bytes, char *err, size_t errlen)
{
    ffr_ctx tmp;
    std::string msg;
    auto fail = [&](const std::string &m, int rc)
    {
        if (err && errlen)
            snprintf(err,errlen,"%s",m.c_str());
        return rc;
    };
    const bool f32 = desc && desc->elem_size == 4;
    if (!(f32 ? pack_blob<float>(&tmp,desc,msg) : pack_blob<double>(&tmp,desc,msg)))
        return fail(msg,FFR_E_INVALID);
    if (!jit_supported(&tmp))
        return fail("the flame-specialised kernel supports <= 8 xforms and <= 4 colour dimensions",FFR_E_UNSUPPORTED);
    const bool ok = jit_prepare(&tmp);
    if (source && source_len)
        snprintf(source,source_len,"%s",tmp.jit_source.c_str());
    if (!ok)
        return fail(tmp.jit_err,FFR_E_UNSUPPORTED);
    if (cubin_bytes)
        *cubin_bytes = tmp.jit_cubin.size();
    return FFR_OK;
}

int ffr_cuda_atomic_roofline(ffr_ctx *ctx, uint64_t n_atomics, int pattern, float *ms)
{
    return ffr_cuda_atomic_roofline_ex(ctx,n_atomics,pattern,ms,nullptr);
}

int ffr_cuda_atomic_roofline_ex(ffr_ctx *ctx, uint64_t n_atomics, int pattern, float *ms,
        uint64_t *n_done)
{
    if (!ctx || !ms)
        return FFR_E_INVALID;
    if (pattern < 0 || pattern > 2)
    {
        ctx->err = "atomic_roofline: pattern must be 0 (uniform cells), 1 (attractor replay, streamed trace) "
                   "or 2 (attractor replay from shared-memory windows)";
        return FFR_E_INVALID;
    }
    DeviceState &ds = ctx->devs[0];
    CK(cudaSetDevice(ds.dev));
    /* both microbenchmarks: every SM full (8 x 256 threads), REDs issued back to back */
    const u64 grid = (u64)ds.sm_count * 8;
    EventPair ev;
    CK(cudaEventCreate(&ev.e0));
    CK(cudaEventCreate(&ev.e1));
    if (pattern == 0)
    {
        const u64 threads = grid*FFR_REPLAY_TPB;
        const u64 per_thread = std::max<u64>(8,(n_atomics/threads) & ~7ULL);
        CK(cudaEventRecord(ev.e0,ds.stream));
        if (ctx->elem == 8)
            atomic_bench_kernel<double><<<(unsigned)grid,FFR_REPLAY_TPB,0,ds.stream>>>((u64*)ds.buffer,ctx->cells,
                ctx->cellsz,per_thread,0x1234u + ctx->launches);
        else
            atomic_bench_kernel<float><<<(unsigned)grid,FFR_REPLAY_TPB,0,ds.stream>>>((unsigned int*)ds.buffer,
                ctx->cells,ctx->cellsz,per_thread,0x1234u + ctx->launches);
        ++ctx->launches;
        CK(cudaGetLastError());
        CK(cudaEventRecord(ev.e1,ds.stream));
        CK(cudaEventSynchronize(ev.e1));
        CK(cudaEventElapsedTime(ms,ev.e0,ev.e1));
        if (n_done)
            *n_done = per_thread*threads;
        ds.dirty = true;
        return FFR_OK;
    }
    /* record: the chains the render kernel keeps resident, n_atomics/chains samples each (an even
       number of chains: the replay reads entry pairs); statistics saved and restored */
    u64 chains = std::max<u64>(2,ffr_cuda_resident_chains(ctx)) & ~1ULL;
    /* pattern 2 replays windows many times over: a trace that fills them is enough */
    const u64 want = pattern == 2 ? std::min<u64>(n_atomics,(u64)ds.sm_count*2*(96u*1024u/8u)) : n_atomics;
    const u64 per_chain = std::max<u64>(1,(want + chains - 1)/chains);
    const u64 entries = chains*per_chain;
    DevTemp trace, saved;
    CK(trace.alloc(entries*sizeof(u64)));
    CK(saved.alloc(sizeof(DevStats)));
    CK(cudaMemsetAsync(trace.p,0xff,entries*sizeof(u64),ds.stream));
    CK(cudaMemcpyAsync(saved.p,ds.d_stats,sizeof(DevStats),cudaMemcpyDeviceToDevice,ds.stream));
    const uint32_t mode = ctx->scatter_mode;
    ctx->scatter_mode = FFR_SCATTER_TRACE;
    ds.d_trace = trace.as<u64>();
    int rc = launch_render(ctx,ds,0x7ace0000ULL,chains,per_chain,0,1,~0ULL);
    ctx->scatter_mode = mode;
    ds.d_trace = nullptr;
    if (rc != FFR_OK)
        return rc;
    DevStats after, before;
    CK(cudaMemcpyAsync(&after,ds.d_stats,sizeof(DevStats),cudaMemcpyDeviceToHost,ds.stream));
    CK(cudaMemcpyAsync(&before,saved.p,sizeof(DevStats),cudaMemcpyDeviceToHost,ds.stream));
    CK(cudaStreamSynchronize(ds.stream));
    if (n_done)
        *n_done = after.s_plot - before.s_plot;
    CK(cudaMemcpyAsync(ds.d_stats,saved.p,sizeof(DevStats),cudaMemcpyDeviceToDevice,ds.stream));
    if (pattern == 2)
    {
        /* two blocks of 1024 threads per SM, each replaying its own 96 KiB window of the trace */
        const uint32_t win = 96u*1024u/8u;
        const size_t smem = (size_t)win*8;
        const u64 wgrid = std::min<u64>((u64)ds.sm_count*2,(entries + win - 1)/win);
        const u64 covered = std::min<u64>(entries,wgrid*win);
        const uint32_t reps = (uint32_t)std::max<u64>(1,n_atomics/std::max<u64>(1,covered));
        CK(cudaMemsetAsync(ds.d_scratch,0,sizeof(u64),ds.stream));
        if (ctx->elem == 8)
            CK(cudaFuncSetAttribute((const void*)atomic_replay_window_kernel<double>,
                cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem));
        else
            CK(cudaFuncSetAttribute((const void*)atomic_replay_window_kernel<float>,
                cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem));
        CK(cudaEventRecord(ev.e0,ds.stream));
        if (ctx->elem == 8)
            atomic_replay_window_kernel<double><<<(unsigned)wgrid,1024,smem,ds.stream>>>((u64*)ds.buffer,(u64*)ds.d_acc,
                trace.as<u64>(),entries,win,reps,ctx->cellsz,ds.d_scratch);
        else
            atomic_replay_window_kernel<float><<<(unsigned)wgrid,1024,smem,ds.stream>>>((unsigned int*)ds.buffer,
                (unsigned int*)ds.d_acc,trace.as<u64>(),entries,win,reps,ctx->cellsz,ds.d_scratch);
        ++ctx->launches;
        CK(cudaGetLastError());
        CK(cudaEventRecord(ev.e1,ds.stream));
        u64 n_red = 0;
        CK(cudaMemcpyAsync(&n_red,ds.d_scratch,sizeof(u64),cudaMemcpyDeviceToHost,ds.stream));
        CK(cudaStreamSynchronize(ds.stream));
        if (n_done)
            *n_done = n_red;
    }
    else
    {
        CK(cudaEventRecord(ev.e0,ds.stream));
        if (ctx->elem == 8)
            atomic_replay_kernel<double><<<(unsigned)grid,FFR_REPLAY_TPB,0,ds.stream>>>((u64*)ds.buffer,(u64*)ds.d_acc,
                trace.as<ulonglong2>(),entries/2,ctx->cellsz);
        else
            atomic_replay_kernel<float><<<(unsigned)grid,FFR_REPLAY_TPB,0,ds.stream>>>((unsigned int*)ds.buffer,
                (unsigned int*)ds.d_acc,trace.as<ulonglong2>(),entries/2,ctx->cellsz);
        ++ctx->launches;
        CK(cudaGetLastError());
        CK(cudaEventRecord(ev.e1,ds.stream));
    }
    CK(cudaEventSynchronize(ev.e1));
    CK(cudaEventElapsedTime(ms,ev.e0,ev.e1));
    /* what the replay put into K1e's tile goes where a render's would: the tile stays all zero
       between launches */
    rc = fold_tiles(ctx,ds);
    if (rc != FFR_OK)
        return rc;
    CK(cudaStreamSynchronize(ds.stream));
    ds.dirty = true;
    return FFR_OK;
}

} // extern "C"
