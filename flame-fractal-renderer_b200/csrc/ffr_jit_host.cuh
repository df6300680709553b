/*
ffr_jit_host.cuh -- host side of the flame-specialised kernel K1c (ffr_jit_kernel.cuh):
  1. generate(): turn the packed flame blob (the very bytes the interpreter kernels read) into
     CUDA source in which every xform is a straight-line function with literal coefficients;
     literals are hexadecimal floating point, so the generated code holds exactly the values
     the blob holds;
  2. compile(): NVRTC -> sm_100a cubin (-fmad=false like the ahead-of-time build), cached in
     the process and on disk by a hash of source + headers + compiler version;
  3. Module: load the cubin into a device's primary context and launch it (driver API).
libnvrtc and libcuda are opened with dlopen so that libffr_cuda.so keeps loading on a machine
without them (the CPU test box); a missing library only makes the JIT path unavailable, the
ahead-of-time kernels then serve every flame.
*/

#pragma once

#include <cuda.h>
#include <nvrtc.h>
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <cmath>
#include <cstdlib>
#include <map>
#include <mutex>
#include <sstream>

#include "ffr_params.cuh"

/* the device headers, embedded at build time (Makefile: build/ffr_embed.inc) */
#include "ffr_embed.inc"

namespace jit
{

struct Config
{
    int tpb = 256;          /* threads per block */
    int ns = 512;           /* chain slots per block */
    int cap = 512;          /* K1d: ring capacity, power of two >= ns */
    int minb = 2;           /* blocks per SM asked of ptxas (register cap) */
    bool inline_math = false;
    bool async = true;      /* K1d (ffr_jit_async.cuh, queue scheduled) instead of K1c (lock step) */
    bool affine = false;    /* K1e (ffr_jit_affine.cuh): pure-affine flame, one chain per thread */
    int npair = 0;          /* K1e: 16-byte rows of the per-xform coefficient table */
    unsigned acc_mul = 0;   /* K1e: scramble multiplier of the accumulation tile (0: scatter into the buffer) */
    unsigned acc_gran = 0;  /* K1e: log2 of the cells that stay together in the tile */
    unsigned dir_cap = 0;   /* K1e: rows of the compact tile (0: none); excludes acc_mul */
    unsigned dir_cache_sets = 0;     /* K1e compact tile: sets (2 ways each) of the shared-memory directory cache */
    unsigned dir_blk[3] = {0,0,0};   /* K1e compact tile: log2 of a row's extent along each axis (sum 9: a row is
                                        a BLOCK of 512 cells, e.g. 8x8x8); all zero: 512 consecutive cells */
};

struct Api
{
    bool tried = false, ok = false;
    std::string err;
    decltype(&nvrtcCreateProgram) CreateProgram = nullptr;
    decltype(&nvrtcCompileProgram) CompileProgram = nullptr;
    decltype(&nvrtcDestroyProgram) DestroyProgram = nullptr;
    decltype(&nvrtcGetProgramLogSize) GetProgramLogSize = nullptr;
    decltype(&nvrtcGetProgramLog) GetProgramLog = nullptr;
    decltype(&nvrtcGetCUBINSize) GetCUBINSize = nullptr;
    decltype(&nvrtcGetCUBIN) GetCUBIN = nullptr;
    decltype(&nvrtcVersion) Version = nullptr;
    decltype(&cuModuleLoadData) ModuleLoadData = nullptr;
    decltype(&cuModuleUnload) ModuleUnload = nullptr;
    decltype(&cuModuleGetFunction) ModuleGetFunction = nullptr;
    decltype(&cuFuncSetAttribute) FuncSetAttribute = nullptr;
    decltype(&cuFuncGetAttribute) FuncGetAttribute = nullptr;
    decltype(&cuOccupancyMaxActiveBlocksPerMultiprocessor) Occupancy = nullptr;
    decltype(&cuLaunchKernel) LaunchKernel = nullptr;
    decltype(&cuGetErrorString) GetErrorString = nullptr;
};

inline void *open_first(const char *const *names)
{
    for (; *names; ++names)
        if (void *h = dlopen(*names,RTLD_NOW|RTLD_LOCAL))
            return h;
    return nullptr;
}

/* need_driver = false: only the compiler (source generation + NVRTC work without a GPU) */
inline Api &api(bool need_driver)
{
    static Api a;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (!a.tried)
    {
        a.tried = true;
        /* the toolkit's compiler first (the one the ahead-of-time kernels were built with); a
           host process may have another libnvrtc.so.12 loaded already (PyTorch bundles one) */
        static const char *const rtc[] = {"/usr/local/cuda/lib64/libnvrtc.so.12","/usr/local/cuda/lib64/libnvrtc.so",
            "libnvrtc.so.12","libnvrtc.so",nullptr};
        void *h = open_first(rtc);
        if (!h)
        {
            a.err = "libnvrtc not found";
            return a;
        }
#define RTC_SYM(field,name) a.field = (decltype(a.field))dlsym(h,#name); if (!a.field) { a.err = "libnvrtc: missing " #name; return a; }
        RTC_SYM(CreateProgram,nvrtcCreateProgram)
        RTC_SYM(CompileProgram,nvrtcCompileProgram)
        RTC_SYM(DestroyProgram,nvrtcDestroyProgram)
        RTC_SYM(GetProgramLogSize,nvrtcGetProgramLogSize)
        RTC_SYM(GetProgramLog,nvrtcGetProgramLog)
        RTC_SYM(GetCUBINSize,nvrtcGetCUBINSize)
        RTC_SYM(GetCUBIN,nvrtcGetCUBIN)
        RTC_SYM(Version,nvrtcVersion)
#undef RTC_SYM
        a.ok = true;
    }
    if (a.ok && need_driver && !a.LaunchKernel)
    {
        static const char *const drv[] = {"libcuda.so.1","libcuda.so",nullptr};
        void *h = open_first(drv);
        if (!h)
        {
            a.err = "libcuda not found";
            return a;
        }
#define DRV_SYM(field,name) a.field = (decltype(a.field))dlsym(h,#name); if (!a.field) { a.err = "libcuda: missing " #name; return a; }
        DRV_SYM(ModuleLoadData,cuModuleLoadData)
        DRV_SYM(ModuleUnload,cuModuleUnload)
        DRV_SYM(ModuleGetFunction,cuModuleGetFunction)
        DRV_SYM(FuncSetAttribute,cuFuncSetAttribute)
        DRV_SYM(FuncGetAttribute,cuFuncGetAttribute)
        DRV_SYM(Occupancy,cuOccupancyMaxActiveBlocksPerMultiprocessor)
        DRV_SYM(GetErrorString,cuGetErrorString)
        DRV_SYM(LaunchKernel,cuLaunchKernel)
#undef DRV_SYM
    }
    return a;
}

/* ---- source generation ---- */

inline std::string lit(double v)
{
    char buf[80];
    if (std::isfinite(v))
        snprintf(buf,sizeof(buf),"%a",v);
    else
    {
        unsigned long long b;
        memcpy(&b,&v,8);
        snprintf(buf,sizeof(buf),"__longlong_as_double((long long)0x%llxULL)",b);
    }
    return buf;
}

inline std::string lit(float v)
{
    char buf[80];
    if (std::isfinite(v))
        snprintf(buf,sizeof(buf),"%af",(double)v);
    else
    {
        unsigned int b;
        memcpy(&b,&v,4);
        snprintf(buf,sizeof(buf),"__uint_as_float(0x%xu)",b);
    }
    return buf;
}

/* Flame constants go to a __constant__ table (jc[]) rather than into the instruction stream:
   an fp64 literal costs two move instructions at every use, a constant-bank operand none.
   0 and +-1 stay literals so that the compiler can drop exact no-ops (x*1.0). */
template <typename T> struct Pool
{
    std::vector<T> vals;
    std::string ref(T v)
    {
        if (v == (T)0 || v == (T)1 || v == (T)-1 || !std::isfinite(v))
            return lit(v);
        size_t i = 0;
        for (; i < vals.size(); ++i)
            if (memcmp(&vals[i],&v,sizeof(T)) == 0)
                break;
        if (i == vals.size())
            vals.push_back(v);
        return "jc[" + std::to_string(i) + "]";
    }
};

template <typename T>
void emit_affine(std::ostringstream &o, Pool<T> &pool, int D, const T *A, const T *b, const char *in, const char *out)
{
    o << "    { const T A[" << D*D << "] = {";
    for (int i = 0; i < D*D; ++i)
        o << (i ? "," : "") << pool.ref(A[i]);
    o << "}; const T b[" << D << "] = {";
    for (int i = 0; i < D; ++i)
        o << (i ? "," : "") << pool.ref(b[i]);
    o << "}; affine_apply<T,JD>(A,b," << in << "," << out << "); }\n";
}

/* XForm::applyIteration (types/xform.hpp:211-227) for ONE xform, unrolled: the same statements
   xform_apply_fn (ffr_device.cuh) executes for this xform, in the same order */
template <typename T>
void emit_xform(std::ostringstream &o, Pool<T> &pool, const std::string &name, int D, const DevXFormT<T> &xf,
        const DevVarT<T> *vars)
{
    o << "__device__ __forceinline__ void " << name << "(const JT *pin, JT *pout, RngT<JT> &rng)\n{\n";
    o << "    typedef JT T;\n    T t[JD], v[JD];\n";
    if (D < 3 || (xf.flags & XF_HAS_PRE))
        emit_affine<T>(o,pool,D,xf.pre_A,xf.pre_b,"pin","t");
    else
        o << "    for (int i = 0; i < JD; ++i) t[i] = pin[i];\n";
    o << "    for (int i = 0; i < JD; ++i) v[i] = 0.0;\n";
    o << "    PolarT<T> P; P.r2 = P.r = P.ang = P.sa = P.ca = 0.0;\n";
    if (D == 2 && xf.need != 0)
        o << "    JPOLAR(P," << xf.need << "u,t[0],t[1]);\n";
    for (uint32_t k = xf.var_begin; k < xf.var_begin + xf.var_count; ++k)
    {
        const DevVarT<T> &var = vars[k];
        o << "    {\n        DevVarT<T> var; var.op = " << var.op << "u; var.axis_x = " << var.axis_x
          << "u; var.axis_y = " << var.axis_y << "u; var.need = " << var.need << "u; var.weight = "
          << pool.ref(var.weight) << ";\n       ";
        for (int q = 0; q < FFR_MAX_VAR_PARAMS; ++q)
            o << " var.p[" << q << "] = " << pool.ref(var.p[q]) << ";";
        o << "\n        T c[JD];\n";
        const bool v2d = D >= 2 && var.op >= FFR_VAR_FIRST_2D && var.op <= FFR_VAR_LAST_2D;
        if (var.op == FFR_VAR_LINEAR)
            o << "        for (int i = 0; i < JD; ++i) c[i] = t[i];\n";
        else if (v2d)
        {
            if (D > 2)
            {
                /* VariationFrom2D::calc_h, variations.hpp:94-105 */
                o << "        const T a0 = t[" << var.axis_x << "], a1 = t[" << var.axis_y << "];\n";
                if (var.need != 0)
                    o << "        JPOLAR(P," << var.need << "u,a0,a1);\n";
            }
            else
                o << "        const T a0 = t[0], a1 = t[1];\n";
            o << "        T ox, oy;\n        calc2d_body<T," << var.op << "u>(var,rng,P,a0,a1,ox,oy);\n";
            if (D > 2)
            {
                for (int i = 0; i < D; ++i)
                    o << "        c[" << i << "] = " << ((uint32_t)i == var.axis_x ? "ox" :
                        ((uint32_t)i == var.axis_y ? "oy" : "0.0")) << ";\n";
            }
            else
                o << "        c[0] = ox; c[1] = oy;\n";
        }
        else
        {
            o << "        for (int i = 0; i < JD; ++i) c[i] = 0.0;\n";
            o << "        calc_nd_body<T,JD," << var.op << "u>(var,rng,t,c);\n";
        }
        /* v += weight * calc(t): calc[i]*weight then add (point.hpp:215-225) */
        o << "        for (int i = 0; i < JD; ++i) v[i] += c[i] * var.weight;\n    }\n";
    }
    if (D < 3 || (xf.flags & XF_HAS_POST))
        emit_affine<T>(o,pool,D,xf.post_A,xf.post_b,"v","pout");
    else
        o << "    for (int i = 0; i < JD; ++i) pout[i] = v[i];\n";
    o << "}\n\n";
}

template <typename T>
void emit_blend(std::ostringstream &o, Pool<T> &pool, int R, const DevXFormT<T> &xf, const T *colors,
        const char *src, const char *dst, const char *indent)
{
    /* (1.0 - s)*c + s*colour with s and colour of type num_t (render_iterator.hpp:118-121,128-131) */
    o << indent << "{ const T s = " << pool.ref(xf.color_speed) << ";";
    for (int j = 0; j < R; ++j)
        o << " { const T col = " << pool.ref(colors[xf.color_off + j]) << "; " << dst << "[" << j
          << "] = (1.0-s)*" << src << "[" << j << "] + s*col; }";
    o << " }\n";
}

template <typename T>
std::string generate(const std::vector<unsigned char> &blobv, const std::vector<unsigned char> &colorsv,
        const u64 *m0, const unsigned int *m0_32, const Config &cfg)
{
    const unsigned char *blob = blobv.data();
    const DevFlameT<T> *fl = (const DevFlameT<T>*)blob;
    const DevXFormT<T> *xfs = (const DevXFormT<T>*)(blob + fl->xf_off);
    const DevVarT<T> *vars = (const DevVarT<T>*)(blob + fl->var_off);
    const T *colors = (const T*)colorsv.data();
    const int D = (int)fl->dims, R = (int)fl->r, NX = (int)fl->num_xforms;
    std::ostringstream h, o;   /* head (needs the finished constant pool) and body */
    Pool<T> pool;
    h << "/* generated by libffr_cuda for one flame (ffr_jit_host.cuh); kernel: " << (cfg.async ? "ffr_jit_async.cuh" : "ffr_jit_kernel.cuh") << " */\n";
    h << "#define FFR_TPB " << cfg.ns << "\n";
    if (cfg.inline_math)
        h << "#define FFR_MATH_ATTR __forceinline__\n";
    h << "#define FFR_ISAAC_M0_INIT {";
    for (int i = 0; i < 16; ++i)
        h << (i ? "," : "") << m0[i] << "ULL";
    h << "}\n#define FFR_ISAAC_M0_32_INIT {";
    for (int i = 0; i < 16; ++i)
        h << (i ? "," : "") << m0_32[i] << "u";
    h << "}\n";
    if (getenv("FFR_SC_XOR") && *getenv("FFR_SC_XOR") == '0')
        h << "#define FFR_SC_XOR 0\n";
    /* K1d: warps of one scheduler scan the queues from the same position (measured +1.6 %,
       3 of 3 A/B pairs: fewer xform bodies per instruction cache; starting every scan at the queue
       just served instead measured -0.4 %); FFR_JIT_ROT_STATIC=0 rotates */
    if (!(getenv("FFR_JIT_ROT_STATIC") && *getenv("FFR_JIT_ROT_STATIC") == '0'))
        h << "#define JROT_STATIC 1\n";
    /* experiment switch (never default): let ptxas contract a*b+c into FMAs. Breaks the exactness
       contract, so only meaningful for flames that are compared statistically anyway */
    if (getenv("FFR_JIT_FMAD") && *getenv("FFR_JIT_FMAD") == '1')
        h << "/*FFR_NVRTC_FMAD*/\n";
    if (cfg.async)
    {
        /* cold paths out of line: bit 0 chain start, 1 idle path, 2 bad-value record. Measured in all
           eight combinations (csci6360_project / tkoz_test3 at 4096^2): only the bad-value record is
           worth it (2.662e10 -> 2.671e10, 2.589e10 -> 2.601e10); chain start and idle path out of line
           cost up to 8 % (their call sites keep more of the loop's state alive across a call) */
        const char *c = getenv("FFR_JIT_COLD");
        const int m = c ? atoi(c) : 4;
        h << "#define JCOLD_START " << (m & 1) << "\n#define JCOLD_IDLE " << ((m >> 1) & 1) << "\n#define JCOLD_BAD " << ((m >> 2) & 1) << "\n";
    }
    if (cfg.async && !(getenv("FFR_JIT_SIN_VIA_SINCOS") && *getenv("FFR_JIT_SIN_VIA_SINCOS") == '0'))
        h << "#define FFR_SIN_VIA_SINCOS 1\n";
    /* K1d: gen() rolled to the four-step pattern of rngstep4 (a quarter of the unrolled code), or,
       for flames whose xform bodies are large, to ONE step in the loop body (2 KB less hot code for
       7 instructions more per iteration): csci6360_project (19 variations) 2.559e10 -> 2.658e10
       with the one-step form, tkoz_test3 (14 variations) 2.583e10 -> 2.528e10. The threshold between
       the two is the number of variations in the flame. FFR_JIT_GEN_ROLLED=0/1/2 forces a form. */
    if (cfg.async)
    {
        uint32_t nvar = 0;
        for (int k = 0; k < NX + (fl->has_final ? 1 : 0); ++k)
            nvar += xfs[k].var_count;
        int form = (nvar >= 16 && sizeof(T) == 8) ? 2 : 1;   /* the float build's code is small enough either way (3.53e10 vs 3.58e10 with the four-step form) */
        if (const char *e = getenv("FFR_JIT_GEN_ROLLED"))
            form = atoi(e);
        if (form)
            h << "#define FFR_GEN_ROLLED " << form << "\n";
    }
    /* K1d: queue entries as release stores / acquire loads (default); FFR_JIT_ACQREL=0 compiles the
       round-1 form (volatile accesses, no MEMBAR.CTA) for A/B measurements */
    if (cfg.async && getenv("FFR_JIT_ACQREL") && *getenv("FFR_JIT_ACQREL") == '0')
        h << "#define JQ_ACQREL 0\n";
    if (cfg.async)
        h << "#define JRSL_SMEM " << (fl->uses_rng ? 1 : 0) << "\n"
          << (fl->uses_rng ? "" : "#define FFR_RSL_LOAD(p) __ldcg(p)\n");
    /* K1d, the out-of-line arithmetic units and the instruction cache. The hot code of a
       variation flame is about as large as the 32 KB instruction cache, and what fits decides more
       than the instruction count does (measured, csci6360_project / tkoz_test3 at 4096^2, samples/s):
         general polar unit, separate sin / cos / sincos units      2.544e10 / 2.330e10
         polar unit per need mask (54 instead of 75 instructions,
           one reciprocal for both divisions; 1-3 copies)            2.503e10 / 2.426e10
         sin and cos through the sincos unit (two units less)        2.490e10 / 2.338e10
         both                                                        2.562e10 / 2.557e10
       Both are the default; FFR_JIT_POLAR_NEED=0 / FFR_JIT_SIN_VIA_SINCOS=0 compile the other forms.
       With atan2 called from the xform body instead of from inside the polar unit (copies shared
       across NEED_ANG, a unit without calls): 2.511e10 / 2.583e10, and together with the one-step
       gen() below 2.658e10 / 2.528e10. */
    bool polar_per_mask = cfg.async;
    if (const char *e = getenv("FFR_JIT_POLAR_NEED"))
        polar_per_mask = cfg.async && *e != '0';
    if (cfg.async)
        h << ((getenv("FFR_JIT_SC_INLINE") && *getenv("FFR_JIT_SC_INLINE") == '1') ? "" : "#define FFR_SINCOS_OOL 1\n")
          << (polar_per_mask
                ? ((getenv("FFR_JIT_POLAR_ANG") && *getenv("FFR_JIT_POLAR_ANG") == '0')
                   ? "#define JPOLAR(P,need,x,y) P = polar_fill_need<JT,need>(x,y)\n"
                   /* atan2 called from the xform body, not from inside the polar unit: masks that differ
                      only in NEED_ANG share a copy, and the unit calls nothing on its usual path */
                   : "#define JPOLAR(P,need,x,y) do { if (((need) & ~NEED_ANG) != 0u) P = polar_fill_need<JT,((need) & ~NEED_ANG)>(x,y); "
                     "if ((need) & NEED_ANG) P.ang = m_atan2(y,x); } while (0)\n")
                : "#define JPOLAR(P,need,x,y) P = polar_fill_ool<JT>(need,x,y)\n");
    else
        h << "#define JPOLAR(P,need,x,y) polar_fill(P,need,x,y)\n";
    h << "#include \"ffr_params.cuh\"\n";
    h << "typedef " << (sizeof(T) == 8 ? "double" : "float") << " JT;\n";
    h << "#define JD " << D << "\n#define JR " << R << "\n#define JNX " << NX << "\n#define JNS " << cfg.ns
      << "\n#define JCAP " << cfg.cap
      << "\n#define JTPB " << cfg.tpb << "\n#define JMINB " << cfg.minb << "\n#define JHAS_FINAL "
      << (fl->has_final ? 1 : 0) << "\n#define JANY_RNG " << (fl->uses_rng ? 1 : 0) << "\n\n";
    for (int k = 0; k < NX + (fl->has_final ? 1 : 0); ++k)
        emit_xform<T>(o,pool,"jx_" + std::to_string(k),D,xfs[k],vars);
    /* xform dispatch (the index is warp-uniform in K1c's hot loop) */
    o << "__device__ __forceinline__ void jit_xform(unsigned k, const JT *pin, JT *pout, RngT<JT> &rng)\n{\n    switch (k)\n    {\n";
    for (int k = 0; k < NX; ++k)
    {
        if (k + 1 < NX) o << "    case " << k << ":";
        else o << "    default:";
        o << " jx_" << k << "(pin,pout,rng); break;\n";
    }
    o << "    }\n}\n\n";
    o << "__device__ __forceinline__ void jit_final(const JT *pin, JT *pout, RngT<JT> &rng)\n{\n";
    if (fl->has_final)
        o << "    jx_" << NX << "(pin,pout,rng);\n";
    else
        o << "    for (int i = 0; i < JD; ++i) pout[i] = pin[i];\n";
    o << "}\n\n";
    /* Flame::getRandomXForm (types/flame.hpp:212-219): number of table entries below r */
    o << "__device__ __forceinline__ unsigned jit_select(JT r)\n{\n    unsigned i = 0;\n";
    for (int k = 0; k + 1 < NX; ++k)
        o << "    i += (" << pool.ref(fl->xfcw[k]) << " < r) ? 1u : 0u;\n";
    o << "    return i;\n}\n\n";
    /* the same selection straight from a generator word: randNum() = m / 2^b with m = word >> s
       (b = 53, s = 11 for u64; b = 24, s = 8 for u32, flame_rng.hpp:67-87), and xfcw[k] < m / 2^b
       <=> m > floor(xfcw[k] * 2^b) for integer m (scaling by 2^b is exact), so the integer compare
       selects exactly what the floating point compare selects */
    o << "__device__ __forceinline__ unsigned jit_select_word(Real<JT>::word w)\n{\n    unsigned i = 0;\n"
      << "    const Real<JT>::word m = w >> " << (sizeof(T) == 8 ? 11 : 8) << ";\n";
    for (int k = 0; k + 1 < NX; ++k)
    {
        const double scaled = std::floor(std::ldexp((double)fl->xfcw[k],sizeof(T) == 8 ? 53 : 24));
        const unsigned long long thr = scaled >= 18446744073709549568.0 ? ~0ULL :
            (scaled <= 0.0 ? 0ULL : (unsigned long long)scaled);
        if (fl->xfcw[k] < (T)0)
            o << "    i += 1u;\n";     /* a negative table entry is below every r */
        else
            o << "    i += (m > (Real<JT>::word)" << thr << "ULL) ? 1u : 0u;\n";
    }
    o << "    return i;\n}\n\n";
    o << "__device__ __forceinline__ bool jit_inb(const JT *pf)\n{\n    bool inb = true;\n";
    for (int i = 0; i < D; ++i)
        o << "    inb &= (pf[" << i << "] >= " << pool.ref(fl->lo[i]) << ") && (pf[" << i << "] <= " << pool.ref(fl->hi[i]) << ");\n";
    o << "    return inb;\n}\n\n";
    /* buffer_renderer.hpp:202-209 */
    o << "__device__ __forceinline__ u64 jit_index(const JT *pf)\n{\n";
    o << "    u64 bi = to_index((pf[0] - " << pool.ref(fl->lo[0]) << ") * " << pool.ref(fl->mult_d[0]) << ");\n";
    for (int i = 1; i < D; ++i)
        o << "    bi += to_index((pf[" << i << "] - " << pool.ref(fl->lo[i]) << ") * " << pool.ref(fl->mult_d[i]) << ") * "
          << fl->mult_i[i] << "ULL;\n";
    o << "    return bi;\n}\n\n";
    o << "__device__ __forceinline__ void jit_color(unsigned k, JT *c)\n{\n    typedef JT T;\n    switch (k)\n    {\n";
    for (int k = 0; k < NX; ++k)
        if (R > 0 && (xfs[k].flags & XF_HAS_COLOR))
        {
            o << "    case " << k << ":\n";
            emit_blend<T>(o,pool,R,xfs[k],colors,"c","c","        ");
            o << "        break;\n";
        }
    o << "    default: break;\n    }\n}\n\n";
    o << "__device__ __forceinline__ void jit_final_color(const JT *c, JT *cf)\n{\n    typedef JT T;\n";
    if (R > 0 && fl->has_final && (xfs[NX].flags & XF_HAS_COLOR))
        emit_blend<T>(o,pool,R,xfs[NX],colors,"c","cf","    ");
    else
        o << "    for (int i = 0; i < JR; ++i) cf[i] = c[i];\n";
    o << "}\n\n";
    o << "__device__ __forceinline__ u64 jit_json_id(unsigned k)\n{\n    switch (k)\n    {\n";
    for (int k = 0; k < NX; ++k)
        o << "    case " << k << ": return " << xfs[k].json_id << "ULL;\n";
    o << "    default: return 0;\n    }\n}\n\n";
    o << "#include \"" << (cfg.async ? "ffr_jit_async.cuh" : "ffr_jit_kernel.cuh") << "\"\n";
    h << "__constant__ JT jc[" << (pool.vals.empty() ? 1 : pool.vals.size()) << "] = {";
    for (size_t i = 0; i < pool.vals.size(); ++i)
        h << (i ? "," : "") << lit(pool.vals[i]);
    if (pool.vals.empty())
        h << "0";
    h << "};\n\n";
    return h.str() + o.str();
}

/* ---- K1e: pure-affine flames (ffr_jit_affine.cuh) ----

   The xform of a pure-affine flame is (types/xform.hpp:211-227, types/affine.hpp:104-110,
   types/point.hpp:215-225), per output coordinate i:
       t[i]   = b_pre[i]  + (((0 + A_pre[i][0]*p[0]) + A_pre[i][1]*p[1]) [+ A_pre[i][2]*p[2]])
       v[i]   = (0 + t[i]*w_0) [+ t[i]*w_1 ...]          (one term per `linear` variation)
       out[i] = b_post[i] + (((0 + A_post[i][0]*v[0]) + ...))
   Every xform runs this same code on its own coefficients, so the generated function takes a
   coefficient either from the constant bank (identical for all xforms) or from a per-xform
   table in shared memory (it differs). Coefficients known at compile time allow steps to be
   dropped -- but only where the result is PROVABLY the same IEEE number, sign of zero included
   (there is no -ffast-math anywhere). With every intermediate finite (checked below from the
   coefficient magnitudes) the rules are:
     R1  x*1 == x,  x*(-1) == -x   exactly, zeros included.
     R2  a*x with a == +-0 is +-0, and z + (+-0) == z for z != 0: a zero term changes at most
         the SIGN OF A ZERO partial sum, never a nonzero value. The same holds for the leading
         `0 +`.
     R3  b + S, with b not -0.0: equals b for S == +0 and for S == -0 alike (b != 0: exact;
         b == +0: +0 + (+-0) == +0). So a row whose offset b is never -0.0 ("absorbing") yields
         the reference's bits from ANY S that equals the reference's S as a real number: zero
         terms and the leading 0 are dropped (R2). Its result is never -0.0.
     R4  +0 + S == S when S is never -0.0; a sum is never -0.0 if one operand never is.
   A row whose offset is -0.0 for some xform is emitted in the reference's full form. The sum v
   may drop its leading `0 +` only if every consumer is an absorbing row; otherwise it keeps it
   (or is t itself: one variation of weight 1 and t never -0.0).
   Returns "" and sets `why` if the flame is outside what K1e covers; the caller then keeps the
   interpreter kernel. */
template <typename T> struct AfCoef
{
    bool uniform = true;      /* identical bits for every xform */
    T value = 0;              /* if uniform */
    bool any_negzero = false; /* some xform holds -0.0 here */
    std::string ref;          /* how the generated code reads it */
};

struct AfVal { std::string e; bool nz; };   /* expression, "never -0.0" */

template <typename T>
AfVal af_row(const AfCoef<T> &b, const std::vector<AfCoef<T>> &a, const std::vector<AfVal> &x)
{
    const size_t D = x.size();
    if (b.any_negzero)
    {
        /* the reference's expression in full */
        std::string s = "(T)0.0";
        for (size_t j = 0; j < D; ++j)
            s = "(" + s + " + " + a[j].ref + "*" + x[j].e + ")";
        return {"(" + b.ref + " + " + s + ")",false};
    }
    std::string s;
    bool snz = false;
    for (size_t j = 0; j < D; ++j)
    {
        std::string term;
        bool tnz = false;
        if (a[j].uniform && a[j].value == (T)0)
            continue;                                          /* R2 */
        if (a[j].uniform && a[j].value == (T)1)
        {
            term = x[j].e;                                     /* R1 */
            tnz = x[j].nz;
        }
        else if (a[j].uniform && a[j].value == (T)-1)
            term = "(-" + x[j].e + ")";                        /* R1 */
        else
            term = "(" + a[j].ref + "*" + x[j].e + ")";
        if (s.empty())
        {
            s = term;
            snz = tnz;
        }
        else
        {
            s = "(" + s + " + " + term + ")";
            snz = snz || tnz;
        }
    }
    if (s.empty())
        return {b.ref,true};                                   /* b + (+0) */
    if (b.uniform && b.value == (T)0)
        return {snz ? s : "((T)0.0 + " + s + ")",true};         /* R4 */
    return {"(" + b.ref + " + " + s + ")",true};                /* R3 */
}

template <typename T>
std::string generate_affine(const std::vector<unsigned char> &blobv, const u64 *m0, const unsigned int *m0_32,
        const Config &cfg, int *npair_out, std::string &why)
{
    const unsigned char *blob = blobv.data();
    const DevFlameT<T> *fl = (const DevFlameT<T>*)blob;
    const DevXFormT<T> *xfs = (const DevXFormT<T>*)(blob + fl->xf_off);
    const DevVarT<T> *vars = (const DevVarT<T>*)(blob + fl->var_off);
    const int D = (int)fl->dims, NX = (int)fl->num_xforms;
    auto no = [&](const char *m) { why = std::string("K1e: ") + m; return std::string(); };
    if (fl->r != 0 || fl->has_final)
        return no("colour dimensions or a final xform");
    if (NX < 1 || NX > 8)
        return no("more than 8 xforms");
    const uint32_t nvar = xfs[0].var_count;
    const uint32_t pre = D < 3 ? 1u : (xfs[0].flags & XF_HAS_PRE), post = D < 3 ? 1u : (xfs[0].flags & XF_HAS_POST);
    if (nvar < 1 || nvar > 4)
        return no("0 or more than 4 variations per xform");
    double row_pre = 1.0, row_post = 1.0, b_pre = 0.0, b_post = 0.0, wsum = 0.0;
    for (int k = 0; k < NX; ++k)
    {
        const DevXFormT<T> &xf = xfs[k];
        if (xf.var_count != nvar || (D == 3 && ((xf.flags & XF_HAS_PRE) != pre || (xf.flags & XF_HAS_POST) != post)))
            return no("xforms of different shape");
        double ws = 0.0;
        for (uint32_t q = 0; q < nvar; ++q)
        {
            if (vars[xf.var_begin + q].op != FFR_VAR_LINEAR)
                return no("a variation other than linear");
            ws += std::fabs((double)vars[xf.var_begin + q].weight);
        }
        wsum = std::max(wsum,ws);
        for (int i = 0; i < D; ++i)
        {
            double rp = 0.0, rq = 0.0;
            for (int j = 0; j < D; ++j)
            {
                rp += std::fabs((double)xf.pre_A[i*D + j]);
                rq += std::fabs((double)xf.post_A[i*D + j]);
            }
            if (pre) { row_pre = std::max(row_pre,rp); b_pre = std::max(b_pre,std::fabs((double)xf.pre_b[i])); }
            if (post) { row_post = std::max(row_post,rq); b_post = std::max(b_post,std::fabs((double)xf.post_b[i])); }
        }
    }
    /* every intermediate stays finite: |out| <= alpha*|p| + beta per application, |p| <= 1 at the
       start of a chain, and the settle iterations run unchecked (render_iterator.hpp:55-57) */
    const double alpha = row_pre*wsum*row_post, beta = b_pre*wsum*row_post + b_post;
    const int settle = sizeof(T) == 8 ? 53 : 24;
    const double thr = sizeof(T) == 8 ? 1e20 : 1e10, room = sizeof(T) == 8 ? 290.0 : 34.0;
    if (!(std::isfinite(alpha) && std::isfinite(beta)) ||
            std::log10(std::max(alpha,1.0))*(settle + 2) + std::log10(1.0 + beta) + std::log10(thr)*0.0 + 4.0 > room ||
            std::log10(std::max(alpha,1.0)) + std::log10(thr) + std::log10(1.0 + beta) + 4.0 > room)
        return no("coefficients too large to prove every intermediate finite");
    u64 cells = 1;
    bool idx32 = true;
    for (int i = 0; i < D; ++i)
    {
        /* in bounds => not a bad value (the kernel tests bad values only for unplotted samples) */
        if (!(std::fabs((double)fl->lo[i]) <= thr && std::fabs((double)fl->hi[i]) <= thr))
            return no("bounds beyond the bad value threshold");
        const u64 size_i = (i + 1 < D) ? fl->mult_i[i + 1]/fl->mult_i[i] : fl->cells/fl->mult_i[i];
        if (size_i >= (1ULL << 31))
            idx32 = false;
        cells *= size_i;
    }
    if (fl->cells >= (1ULL << 32))
        idx32 = false;

    /* classify every coefficient position over the xforms */
    Pool<T> pool;
    std::vector<T> table;   /* varying coefficients: [position][xform] */
    auto classify = [&](auto get) {
        AfCoef<T> c;
        c.value = get(0);
        for (int k = 0; k < NX; ++k)
        {
            const T v = get(k);
            if (memcmp(&v,&c.value,sizeof(T)) != 0)
                c.uniform = false;
            if (v == (T)0 && std::signbit(v))
                c.any_negzero = true;
        }
        if (c.uniform)
            c.ref = pool.ref(c.value);
        else
        {
            const size_t pos = table.size()/NX;
            for (int k = 0; k < NX; ++k)
                table.push_back(get(k));
            c.ref = "q" + std::to_string(pos/2) + (pos % 2 ? ".y" : ".x");
        }
        return c;
    };
    std::vector<std::vector<AfCoef<T>>> Apre(D), Apost(D);
    std::vector<AfCoef<T>> bpre(D), bpost(D), w(nvar);
    for (int i = 0; i < D; ++i)
    {
        if (pre)
        {
            for (int j = 0; j < D; ++j)
                Apre[i].push_back(classify([&](int k) { return xfs[k].pre_A[i*D + j]; }));
            bpre[i] = classify([&](int k) { return xfs[k].pre_b[i]; });
        }
    }
    for (uint32_t q = 0; q < nvar; ++q)
        w[q] = classify([&](int k) { return vars[xfs[k].var_begin + q].weight; });
    bool post_absorbing = post != 0;
    for (int i = 0; i < D; ++i)
    {
        if (post)
        {
            for (int j = 0; j < D; ++j)
                Apost[i].push_back(classify([&](int k) { return xfs[k].post_A[i*D + j]; }));
            bpost[i] = classify([&](int k) { return xfs[k].post_b[i]; });
            if (bpost[i].any_negzero)
                post_absorbing = false;
        }
    }
    if (table.size()/NX % 2)
        for (int k = 0; k < NX; ++k)
            table.push_back((T)0);
    const int npair = (int)(table.size()/NX/2);
    if (npair_out) *npair_out = npair;

    /* the expressions; p is never -0.0 (2u-1 is not, and neither is any `out` below) unless a
       non-absorbing row produces it, in which case nothing is assumed about p */
    std::vector<AfVal> out;
    for (int pass = 0; pass < 2; ++pass)
    {
        const bool p_nz = pass == 0;
        std::vector<AfVal> x(D), t(D), v(D);
        for (int i = 0; i < D; ++i)
            x[i] = {"x" + std::to_string(i),p_nz};
        for (int i = 0; i < D; ++i)
            t[i] = pre ? af_row<T>(bpre[i],Apre[i],x) : x[i];
        /* name the rows so that each is evaluated once */
        std::vector<AfVal> tn(D);
        for (int i = 0; i < D; ++i)
            tn[i] = {"t" + std::to_string(i),t[i].nz};
        for (int i = 0; i < D; ++i)
        {
            std::string s;
            bool snz = false;
            for (uint32_t q = 0; q < nvar; ++q)
            {
                std::string term;
                bool tnz = false;
                if (w[q].uniform && w[q].value == (T)1) { term = tn[i].e; tnz = tn[i].nz; }   /* R1 */
                else if (w[q].uniform && w[q].value == (T)-1) term = "(-" + tn[i].e + ")";
                else term = "(" + tn[i].e + "*" + w[q].ref + ")";
                if (s.empty()) { s = term; snz = tnz; }
                else { s = "(" + s + " + " + term + ")"; snz = snz || tnz; }
            }
            if (snz || post_absorbing)
                v[i] = {s,snz};                       /* R4 / consumers are absorbing rows (R3) */
            else
                v[i] = {"((T)0.0 + " + s + ")",true};
        }
        std::vector<AfVal> vn(D);
        for (int i = 0; i < D; ++i)
            vn[i] = {"v" + std::to_string(i),v[i].nz};
        out.assign(D,AfVal());
        bool all_nz = true;
        for (int i = 0; i < D; ++i)
        {
            out[i] = post ? af_row<T>(bpost[i],Apost[i],vn) : vn[i];
            all_nz = all_nz && out[i].nz;
        }
        if (pass == 0 && !all_nz)
            continue;        /* p may be -0.0: redo without that assumption */
        std::ostringstream h, o;
        h << "/* generated by libffr_cuda for one pure-affine flame (ffr_jit_host.cuh); kernel: ffr_jit_affine.cuh */\n";
        h << "#define FFR_TPB " << cfg.tpb << "\n";
        h << "#define FFR_ISAAC_M0_INIT {";
        for (int i = 0; i < 16; ++i)
            h << (i ? "," : "") << m0[i] << "ULL";
        h << "}\n#define FFR_ISAAC_M0_32_INIT {";
        for (int i = 0; i < 16; ++i)
            h << (i ? "," : "") << m0_32[i] << "u";
        h << "}\n#include \"ffr_params.cuh\"\n";
        h << "typedef " << (sizeof(T) == 8 ? "double" : "float") << " JT;\n";
        h << "typedef " << (sizeof(T) == 8 ? "double2" : "float2") << " JPAIR;\n";
        h << "typedef " << (idx32 ? "unsigned int" : "unsigned long long") << " JIDX;\n";
        h << "#define JD " << D << "\n#define JR 0\n#define JNX " << NX << "\n#define JNS " << cfg.tpb
          << "\n#define JTPB " << cfg.tpb << "\n#define JMINB " << cfg.minb << "\n#define JNPAIR " << npair << "\n";
        if (cfg.acc_mul)
            h << "#define JACC_MUL " << cfg.acc_mul << "u\n#define JACC_MASK " << ((fl->cells - 1) >> cfg.acc_gran)
              << "u\n#define JACC_GRAN " << cfg.acc_gran << "\n#define JACC_ELEMS " << fl->cells << "ULL\n";
        if (cfg.dir_cap)
            h << "#define JDIR_CAP " << cfg.dir_cap << "u\n#define JACC_ELEMS "
              << ((unsigned long long)cfg.dir_cap << FFR_DIR_ROW_SHIFT) << "ULL\n";
        if (cfg.dir_cap && (cfg.dir_blk[0] | cfg.dir_blk[1] | cfg.dir_blk[2]))
            h << "#define JDIR_BLOCKED 1\n";
        if (cfg.dir_cap && cfg.dir_cache_sets)
        {
            unsigned bits = 0;
            while ((1u << bits) < cfg.dir_cache_sets)
                ++bits;
            h << "#define JDC_SETS " << cfg.dir_cache_sets << "u\n#define JDC_SETBITS " << bits << "\n";
        }
        h << "\n";
        o << "/* XForm::applyIteration for every xform of the flame; tb = coefficient table + xform index */\n";
        if (p_nz)
            o << "/* relies on: no coordinate of pin is -0.0 (true for 2u-1 and for every pout of this function) */\n";
        o << "__device__ __forceinline__ void jaf_xform(const JPAIR *tb, const JT *pin, JT *pout)\n{\n    typedef JT T;\n";
        for (int r = 0; r < npair; ++r)
            o << "    const JPAIR q" << r << " = tb[" << r << "*JNX];\n";
        for (int i = 0; i < D; ++i)
            o << "    const T x" << i << " = pin[" << i << "];\n";
        for (int i = 0; i < D; ++i)
            o << "    const T t" << i << " = " << t[i].e << ";\n";
        for (int i = 0; i < D; ++i)
            o << "    const T v" << i << " = " << v[i].e << ";\n";
        for (int i = 0; i < D; ++i)
            o << "    pout[" << i << "] = " << out[i].e << ";\n";
        o << "}\n\n";
        /* Flame::getRandomXForm (types/flame.hpp:212-219), as in generate() */
        o << "__device__ __forceinline__ unsigned jit_select(JT r)\n{\n    unsigned i = 0;\n";
        for (int k = 0; k + 1 < NX; ++k)
            o << "    i += (" << pool.ref(fl->xfcw[k]) << " < r) ? 1u : 0u;\n";
        o << "    return i;\n}\n\n";
        /* on the raw word: xfcw[k] < (w >> s)/2^b  <=>  (w >> s) > floor(xfcw[k]*2^b)  <=>
           w > (floor(xfcw[k]*2^b) << s | (2^s - 1)) */
        o << "__device__ __forceinline__ unsigned jit_select_word(Real<JT>::word w)\n{\n    unsigned i = 0;\n";
        for (int k = 0; k + 1 < NX; ++k)
        {
            const int bbits = sizeof(T) == 8 ? 53 : 24, sbits = sizeof(T) == 8 ? 11 : 8;
            const double scaled = std::floor(std::ldexp((double)fl->xfcw[k],bbits));
            if (fl->xfcw[k] < (T)0)
                o << "    i += 1u;\n";
            else if (scaled >= std::ldexp(1.0,bbits))
                o << "    /* table entry >= 1: never below r */\n";
            else
            {
                const unsigned long long m = (unsigned long long)scaled;
                const unsigned long long thrw = (m << sbits) | ((1ULL << sbits) - 1ULL);
                o << "    i += (w > (Real<JT>::word)" << thrw << "ULL) ? 1u : 0u;\n";
            }
        }
        o << "    return i;\n}\n\n";
        o << "__device__ __forceinline__ bool jit_inb(const JT *pf)\n{\n    bool inb = true;\n";
        for (int i = 0; i < D; ++i)
            o << "    inb &= (pf[" << i << "] >= " << pool.ref(fl->lo[i]) << ") & (pf[" << i << "] <= " << pool.ref(fl->hi[i]) << ");\n";
        o << "    return inb;\n}\n\n";
        /* buffer_renderer.hpp:202-209; sizes < 2^31 and cells < 2^32: the truncating conversion
           and the index arithmetic fit 32 bits (same values) */
        const char *cvt = idx32 ? (sizeof(T) == 8 ? "__double2uint_rz" : "__float2uint_rz") : "to_index";
        o << "__device__ __forceinline__ JIDX jaf_index(const JT *pf)\n{\n";
        o << "    JIDX bi = " << cvt << "((pf[0] - " << pool.ref(fl->lo[0]) << ") * " << pool.ref(fl->mult_d[0]) << ");\n";
        for (int i = 1; i < D; ++i)
            o << "    bi += (JIDX)" << cvt << "((pf[" << i << "] - " << pool.ref(fl->lo[i]) << ") * " << pool.ref(fl->mult_d[i]) << ") * (JIDX)"
              << fl->mult_i[i] << "ULL;\n";
        o << "    return bi;\n}\n\n";
        if (cfg.dir_cap && (cfg.dir_blk[0] | cfg.dir_blk[1] | cfg.dir_blk[2]))
        {
            /* compact tile with BLOCK rows: the row a cell belongs to and its place inside, from the
               per-axis cell coordinates; bi (the reference's linear index, :202-209) is only used
               when the row has no slot in the tile */
            o << "__device__ __forceinline__ void jaf_dir_split(const JT *pf, unsigned &row, unsigned &off, JIDX &bi)\n{\n";
            u64 row_mult = 1;
            unsigned off_shift = 0;
            std::string row_e, off_e, bi_e;
            for (int i = 0; i < D; ++i)
            {
                o << "    const unsigned c" << i << " = " << (idx32 ? cvt : (sizeof(T) == 8 ? "(unsigned)__double2ull_rz" : "(unsigned)__float2ull_rz"))
                  << "((pf[" << i << "] - " << pool.ref(fl->lo[i]) << ") * " << pool.ref(fl->mult_d[i]) << ");\n";
                const u64 size_i = (i + 1 < D) ? fl->mult_i[i + 1]/fl->mult_i[i] : fl->cells/fl->mult_i[i];
                const unsigned b = cfg.dir_blk[i];
                row_e += (i ? " + " : "") + std::string("(c") + std::to_string(i) + " >> " + std::to_string(b) + ")*" + std::to_string(row_mult) + "u";
                off_e += (i ? " | " : "") + std::string("((c") + std::to_string(i) + " & " + std::to_string((1u << b) - 1u) + "u) << " + std::to_string(off_shift) + ")";
                bi_e += (i ? " + " : "") + std::string("(JIDX)c") + std::to_string(i) + "*(JIDX)" + std::to_string(fl->mult_i[i]) + "ULL";
                row_mult *= size_i >> b;
                off_shift += b;
            }
            o << "    row = " << row_e << ";\n    off = " << off_e << ";\n    bi = " << bi_e << ";\n}\n\n";
        }
        o << "__device__ __forceinline__ u64 jit_json_id(unsigned k)\n{\n    switch (k)\n    {\n";
        for (int k = 0; k < NX; ++k)
            o << "    case " << k << ": return " << xfs[k].json_id << "ULL;\n";
        o << "    default: return 0;\n    }\n}\n\n";
        o << "#include \"ffr_jit_affine.cuh\"\n";
        h << "__constant__ JT jc[" << (pool.vals.empty() ? 1 : pool.vals.size()) << "] = {";
        for (size_t i = 0; i < pool.vals.size(); ++i)
            h << (i ? "," : "") << lit(pool.vals[i]);
        if (pool.vals.empty())
            h << "0";
        h << "};\n";
        h << "__constant__ JT jtab[" << (table.empty() ? 2 : table.size()) << "] = {";
        /* [pair][xform][2] */
        if (table.empty())
            h << "0,0";
        for (int r = 0; r < npair; ++r)
            for (int k = 0; k < NX; ++k)
                h << ((r || k) ? "," : "") << lit(table[(size_t)(2*r)*NX + k]) << "," << lit(table[(size_t)(2*r + 1)*NX + k]);
        h << "};\n\n";
        return h.str() + o.str();
    }
    return no("internal");
}

/* ---- compile + cache ---- */

inline u64 fnv1a(const void *data, size_t n, u64 h = 0xcbf29ce484222325ULL)
{
    const unsigned char *p = (const unsigned char*)data;
    for (size_t i = 0; i < n; ++i)
    {
        h ^= p[i];
        h *= 0x100000001b3ULL;
    }
    return h;
}

/* a second, independent 64-bit hash of the same bytes (multiply-xorshift over 8-byte words):
   together with fnv1a the cache key is 128 bits, and both halves are stored INSIDE the entry */
inline u64 mxhash(const void *data, size_t n, u64 h = 0x9e3779b97f4a7c15ULL)
{
    const unsigned char *p = (const unsigned char*)data;
    while (n)
    {
        u64 w = 0;
        const size_t k = n < 8 ? n : 8;
        memcpy(&w,p,k);
        p += k;
        n -= k;
        h = (h ^ w ^ ((u64)k << 56)) * 0xff51afd7ed558ccdULL;
        h ^= h >> 32;
        h *= 0xc4ceb9fe1a85ec53ULL;
        h ^= h >> 29;
    }
    return h;
}

/* Where compiled kernels are kept between runs: FFR_JIT_CACHE, else $XDG_CACHE_HOME/ffr-b200-jit,
   else ~/.cache/ffr-b200-jit -- never a shared, predictable path under /tmp. Empty: no disk cache. */
inline std::string cache_dir()
{
    const char *e = getenv("FFR_JIT_CACHE");
    if (e && *e)
        return e;
    e = getenv("XDG_CACHE_HOME");
    if (e && *e == '/')
        return std::string(e) + "/ffr-b200-jit";
    e = getenv("HOME");
    if (e && *e == '/')
        return std::string(e) + "/.cache/ffr-b200-jit";
    return std::string();
}

/* The directory is only used when it is a real directory (not a symlink) that belongs to this
   user and that nobody else can write to or read: a cubin is code that runs in the caller's
   CUDA context. Created 0700 (with its parent) on first use when `create` is set. */
inline bool cache_dir_usable(const std::string &dir, bool create)
{
    if (dir.empty())
        return false;
    struct stat st;
    if (lstat(dir.c_str(),&st) != 0)
    {
        if (!create)
            return false;
        const size_t cut = dir.find_last_of('/');
        if (cut != std::string::npos && cut > 0)
            mkdir(dir.substr(0,cut).c_str(),0700);     /* e.g. ~/.cache; fine if it exists */
        if (mkdir(dir.c_str(),0700) != 0 && errno != EEXIST)
            return false;
        if (lstat(dir.c_str(),&st) != 0)
            return false;
    }
    return S_ISDIR(st.st_mode) && st.st_uid == getuid() && (st.st_mode & 077) == 0;
}

/* cache entry = header | cubin | register spill bytes (8) */
struct CacheHeader
{
    char magic[8];          /* "FFRJIT2\0" */
    u64 key_fnv, key_mx;    /* the full 128-bit key again: a file under the wrong name is refused */
    u64 key_bytes;          /* length of the hashed material (source + headers + compiler version) */
    u64 payload_bytes;      /* cubin + 8 */
    u64 payload_fnv;        /* truncated or damaged entries are refused */
};

struct Shim { const char *name; const char *text; size_t len; };

inline const std::vector<Shim> &headers()
{
    /* NVRTC has no host headers: the few names the device code takes from them */
    static const char shim_int[] =
        "#pragma once\n"
        "typedef signed char int8_t; typedef unsigned char uint8_t; typedef short int16_t;\n"
        "typedef unsigned short uint16_t; typedef int int32_t; typedef unsigned int uint32_t;\n"
        "typedef long long int64_t; typedef unsigned long long uint64_t;\n";
    static const char shim_math[] =
        "#pragma once\n"
        "#define M_PI 3.14159265358979323846\n"
        "#define M_1_PI 0.31830988618379067154\n"
        "#define M_PI_4 0.78539816339744830962\n"
        "#define INFINITY __int_as_float(0x7f800000)\n";
    static const char shim_empty[] = "#pragma once\n";
    static const std::vector<Shim> h = {
        {"ffr_device.cuh",(const char*)ffr_embed_device,ffr_embed_device_len},
        {"ffr_params.cuh",(const char*)ffr_embed_params,ffr_embed_params_len},
        {"ffr_jit_kernel.cuh",(const char*)ffr_embed_jit_kernel,ffr_embed_jit_kernel_len},
        {"ffr_jit_async.cuh",(const char*)ffr_embed_jit_async,ffr_embed_jit_async_len},
        {"ffr_jit_affine.cuh",(const char*)ffr_embed_jit_affine,ffr_embed_jit_affine_len},
        {"../../include/ffr_cuda.h",(const char*)ffr_embed_abi,ffr_embed_abi_len},
        {"cstdint",shim_int,sizeof(shim_int)-1},
        {"stdint.h",shim_int,sizeof(shim_int)-1},
        {"stddef.h",shim_empty,sizeof(shim_empty)-1},
        {"math.h",shim_math,sizeof(shim_math)-1},
    };
    return h;
}

/* source -> sm_100a cubin. Needs no GPU. */
inline bool compile(const std::string &src, std::vector<char> &cubin, std::string &err, double *seconds,
        bool *from_cache, long *spill_bytes, bool cache_only = false)
{
    /* the last 8 bytes of a cached entry hold the kernel's register spill bytes (ptxas -v) */
    static std::map<std::pair<u64,u64>,std::vector<char>> mem_cache;
    auto split = [&](std::vector<char> &blob)
    {
        long long sp = 0;
        if (blob.size() >= 8)
        {
            memcpy(&sp,blob.data() + blob.size() - 8,8);
            blob.resize(blob.size() - 8);
        }
        if (spill_bytes) *spill_bytes = (long)sp;
    };
    static std::mutex mu;
    Api &a = api(false);
    if (seconds) *seconds = 0.0;
    if (from_cache) *from_cache = false;
    if (!a.ok)
    {
        err = a.err;
        return false;
    }
    int vmaj = 0, vmin = 0;
    a.Version(&vmaj,&vmin);
    u64 h = fnv1a(src.data(),src.size());
    u64 h2 = mxhash(src.data(),src.size());
    u64 key_bytes = src.size();
    for (const Shim &s : headers())
    {
        h = fnv1a(s.text,s.len,h);
        h2 = mxhash(s.text,s.len,h2);
        key_bytes += s.len;
    }
    const int ver[2] = {vmaj,vmin};
    h = fnv1a(ver,sizeof(ver),h);
    h2 = mxhash(ver,sizeof(ver),h2);
    key_bytes += sizeof(ver);
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = mem_cache.find(std::make_pair(h,h2));
        if (it != mem_cache.end())
        {
            cubin = it->second;
            split(cubin);
            if (from_cache) *from_cache = true;
            return true;
        }
    }
    char name[64];
    snprintf(name,sizeof(name),"/%016llx%016llx.cubin",(unsigned long long)h,(unsigned long long)h2);
    const std::string dir = cache_dir();
    const std::string path = dir + name;
    const char *nocache = getenv("FFR_JIT_NO_DISK_CACHE");
    const bool disk = !(nocache && *nocache == '1') && !dir.empty();
    if (disk && cache_dir_usable(dir,false))
    {
        /* O_NOFOLLOW: an entry is a regular file of ours, never a link to somewhere else */
        const int fd = open(path.c_str(),O_RDONLY | O_NOFOLLOW | O_CLOEXEC);
        FILE *f = fd >= 0 ? fdopen(fd,"rb") : nullptr;
        if (!f && fd >= 0)
            close(fd);
        if (f)
        {
            CacheHeader hd;
            struct stat st;
            bool ok = fstat(fd,&st) == 0 && S_ISREG(st.st_mode) && st.st_uid == getuid() &&
                fread(&hd,1,sizeof(hd),f) == sizeof(hd) && memcmp(hd.magic,"FFRJIT2",8) == 0 &&
                hd.key_fnv == h && hd.key_mx == h2 && hd.key_bytes == key_bytes &&
                hd.payload_bytes > 8 && hd.payload_bytes < (1ULL << 30) &&
                (u64)st.st_size == sizeof(hd) + hd.payload_bytes;
            if (ok)
            {
                cubin.resize((size_t)hd.payload_bytes);
                ok = fread(cubin.data(),1,cubin.size(),f) == cubin.size() &&
                    fnv1a(cubin.data(),cubin.size()) == hd.payload_fnv;
            }
            fclose(f);
            if (ok)
            {
                std::lock_guard<std::mutex> lock(mu);
                mem_cache[std::make_pair(h,h2)] = cubin;
                split(cubin);
                if (from_cache) *from_cache = true;
                return true;
            }
            cubin.clear();
        }
    }
    if (cache_only)
    {
        err = "not in the kernel cache";
        return false;
    }
    std::vector<const char*> hn, ht;
    std::vector<std::string> texts;
    for (const Shim &s : headers())
        texts.emplace_back(s.text,s.len);     /* NUL-terminated copies */
    /* FFR_JIT_DUMP_DIR (profiling aid): compile from real files so that -lineinfo points at
       paths ncu --import-source can read: <dir>/p/csrc/ffr_flame.cu + headers */
    std::string prog_name = "ffr_flame.cu", inc_opt;
    const char *dump = getenv("FFR_JIT_DUMP_DIR");
    if (dump && *dump)
    {
        const std::string d = dump;
        mkdir(d.c_str(),0755);
        mkdir((d + "/p").c_str(),0755);
        mkdir((d + "/p/csrc").c_str(),0755);
        mkdir((d + "/include").c_str(),0755);
        auto put = [](const std::string &path, const std::string &text)
        {
            if (FILE *f = fopen(path.c_str(),"wb"))
            {
                fwrite(text.data(),1,text.size(),f);
                fclose(f);
            }
        };
        for (size_t i = 0; i < texts.size(); ++i)
        {
            const std::string nm = headers()[i].name;
            put(nm.rfind("../../",0) == 0 ? d + "/" + nm.substr(6) : d + "/p/csrc/" + nm,texts[i]);
        }
        prog_name = d + "/p/csrc/ffr_flame.cu";
        put(prog_name,src);
        inc_opt = "-I" + d + "/p/csrc";
    }
    else
        for (size_t i = 0; i < texts.size(); ++i)
        {
            hn.push_back(headers()[i].name);
            ht.push_back(texts[i].c_str());
        }
    timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC,&t0);
    nvrtcProgram prog = nullptr;
    if (a.CreateProgram(&prog,src.c_str(),prog_name.c_str(),(int)hn.size(),ht.data(),hn.data()) != NVRTC_SUCCESS)
    {
        err = "nvrtcCreateProgram failed";
        return false;
    }
    const bool fmad = src.find("/*FFR_NVRTC_FMAD*/") != std::string::npos;
    const char *opts[] = {"--gpu-architecture=sm_100a",fmad ? "-fmad=true" : "-fmad=false","-std=c++17","-lineinfo","-default-device",
        "--ptxas-options=-v",inc_opt.c_str()};
    const nvrtcResult rc = a.CompileProgram(prog,inc_opt.empty() ? 6 : 7,opts);
    std::string log;
    {
        size_t n = 0;
        a.GetProgramLogSize(prog,&n);
        log.assign(n,'\0');
        if (n) a.GetProgramLog(prog,&log[0]);
    }
    if (getenv("FFR_JIT_LOG"))       /* development aid: the ptxas -v lines */
        fprintf(stderr,"%s\n",log.c_str());
    if (rc != NVRTC_SUCCESS)
    {
        err = "NVRTC: " + log.substr(0,4000);
        a.DestroyProgram(&prog);
        return false;
    }
    size_t n = 0;
    a.GetCUBINSize(prog,&n);
    cubin.resize(n);
    a.GetCUBIN(prog,cubin.data());
    a.DestroyProgram(&prog);
    /* "Function properties for ffr_jit_render[_modes] ... N bytes spill stores" */
    long long spills = 0;
    {
        size_t at = log.find("Function properties for ffr_jit_render");
        if (at != std::string::npos)
        {
            size_t e = log.find("bytes spill stores",at);
            if (e != std::string::npos)
            {
                size_t b = log.rfind(',',e);
                if (b != std::string::npos && b > at)
                    spills = atoll(log.c_str() + b + 1);
            }
        }
    }
    if (spill_bytes) *spill_bytes = (long)spills;
    std::vector<char> entry = cubin;
    entry.resize(cubin.size() + 8);
    memcpy(entry.data() + cubin.size(),&spills,8);
    clock_gettime(CLOCK_MONOTONIC,&t1);
    if (seconds) *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9*(double)(t1.tv_nsec - t0.tv_nsec);
    if (disk && cache_dir_usable(dir,true))
    {
        CacheHeader hd;
        memset(&hd,0,sizeof(hd));
        memcpy(hd.magic,"FFRJIT2",8);
        hd.key_fnv = h;
        hd.key_mx = h2;
        hd.key_bytes = key_bytes;
        hd.payload_bytes = entry.size();
        hd.payload_fnv = fnv1a(entry.data(),entry.size());
        const std::string tmp = path + "." + std::to_string((long)getpid());
        const int fd = open(tmp.c_str(),O_WRONLY | O_CREAT | O_EXCL | O_NOFOLLOW | O_CLOEXEC,0600);
        FILE *f = fd >= 0 ? fdopen(fd,"wb") : nullptr;
        if (!f && fd >= 0)
            close(fd);
        if (f)
        {
            const bool ok = fwrite(&hd,1,sizeof(hd),f) == sizeof(hd) &&
                fwrite(entry.data(),1,entry.size(),f) == entry.size();
            const bool closed = fclose(f) == 0;
            if (!ok || !closed || rename(tmp.c_str(),path.c_str()) != 0)
                unlink(tmp.c_str());
        }
    }
    std::lock_guard<std::mutex> lock(mu);
    mem_cache[std::make_pair(h,h2)] = entry;
    return true;
}

/* a cubin loaded into one device's primary context (which must be current) */
struct Module
{
    CUmodule mod = nullptr;
    CUfunction fn = nullptr;
    int regs = 0;
};

inline std::string cu_err(Api &a, CUresult r)
{
    const char *s = nullptr;
    if (a.GetErrorString) a.GetErrorString(r,&s);
    return s ? s : "unknown driver error";
}

inline bool load(const std::vector<char> &cubin, size_t smem, Module &m, std::string &err,
        const char *entry = "ffr_jit_render")
{
    Api &a = api(true);
    if (!a.LaunchKernel)
    {
        err = a.err;
        return false;
    }
    CUresult r = a.ModuleLoadData(&m.mod,cubin.data());
    if (r != CUDA_SUCCESS)
    {
        err = "cuModuleLoadData: " + cu_err(a,r);
        return false;
    }
    r = a.ModuleGetFunction(&m.fn,m.mod,entry);
    if (r != CUDA_SUCCESS)
    {
        err = "cuModuleGetFunction: " + cu_err(a,r);
        return false;
    }
    r = a.FuncSetAttribute(m.fn,CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES,(int)smem);
    if (r != CUDA_SUCCESS)
    {
        err = "cuFuncSetAttribute(smem): " + cu_err(a,r);
        return false;
    }
    a.FuncGetAttribute(&m.regs,CU_FUNC_ATTRIBUTE_NUM_REGS,m.fn);
    return true;
}

inline void unload(Module &m)
{
    Api &a = api(true);
    if (m.mod && a.ModuleUnload)
        a.ModuleUnload(m.mod);
    m.mod = nullptr;
    m.fn = nullptr;
}

} // namespace jit
