/*
ffr_cuda.h -- C ABI of libffr_cuda: the B200 (sm_100a) chaos-game
iterate-and-accumulate path of tkoz0/flame-fractal-renderer.

The reference has no plugin/FFI interface. The seam this library replaces is the
public surface of BufferRenderer<dims> as used by run_renderer<dims>() in
src/ffr_buf.cpp:146-274 (reference paths are relative to the reference repo):

  reference (C++ template, host threads)              this ABI (plain C, device)
  ---------------------------------------------------------------------------------
  Flame<dims>(json)            types/flame.hpp:91     ffr_flame_from_json  (ffr_flame.h)
  BufferRenderer(flame)        buffer_renderer.hpp:254 ffr_cuda_create
  addBuffer(istream)           buffer_renderer.hpp:419 ffr_cuda_add_buffer
  render(N,threads,batch,bv)   buffer_renderer.hpp:269 ffr_cuda_render
  renderSeeded(N,batch,bv)     buffer_renderer.hpp:349 ffr_cuda_render / _render_chains
  writeBuffer(ostream)         buffer_renderer.hpp:476 ffr_cuda_read_buffer
  getSamplesIterated and co.   buffer_renderer.hpp:534 ffr_stats
  ~BufferRenderer              (RAII)                  ffr_cuda_destroy

Plain pointers and sizes only; no exceptions cross this boundary; errors are
return codes plus a message (ffr_cuda_last_error / the err buffer of create).

Determinism contract: a render of N samples with chain
length L is the sum over chains k = 0 .. ceil(N/L)-1 of exactly what the
reference computes for
    rng::setSeed((u64) ffr_chain_seed(base_seed, k));
    renderer.renderSeeded(len_k, len_k, bv_limit);          // one _render_batch
with len_k = L except for a short last chain. Counts are bit-exact and independent
of the number of GPUs; colour sums differ only by floating point summation order.
*/

#ifndef FFR_CUDA_H
#define FFR_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FFR_MAX_DIMS        3     /* ffr_buf.cpp:134-142 instantiates dims 1,2,3 */
#define FFR_MAX_COLOR_DIMS  127   /* flame.hpp:165 */
#define FFR_MAX_XFORMS      64    /* device limit (reference: unbounded) */
#define FFR_MAX_VAR_PARAMS  8     /* largest parameter block: mobius, variations.hpp:1607 */
#define FFR_MAX_BAD_RECORDED 1024 /* bad value records kept per render call */
#define FFR_FINAL_XFORM_ID  (~(uint64_t)0) /* flame.hpp:196 passes id -1 */

/* Variation opcodes. Numbering follows the reference factory / SURVEY appendix A.
   Each comment gives the reference class (variations/variations.hpp:line) and the
   layout of ffr_variation.params: the DERIVED members the reference constructor
   stores, computed by the host with the same expressions (ffr_flame_from_json). For
   N-d variations with per-axis vectors, component i lives at params[i] and the
   second vector at params[4+i]. */
enum ffr_var_op
{
    FFR_VAR_LINEAR = 1,      /* Linear :170        -                                   */
    FFR_VAR_SINUSOIDAL,      /* Sinusoidal :184    -                                   */
    FFR_VAR_SPHERICAL,       /* Spherical :198     -                                   */
    FFR_VAR_BENT,            /* Bent :216          [i]=scales_neg [4+i]=scales_pos     */
    FFR_VAR_RECTANGLES,      /* Rectangles :245    [i]=params                          */
    FFR_VAR_FISHEYE,         /* Fisheye :275       [0]=addval                          */
    FFR_VAR_BUBBLE,          /* Bubble :297        [0]=addval                          */
    FFR_VAR_NOISE,           /* Noise :319         -                                   */
    FFR_VAR_BLUR,            /* Blur :335          -                                   */
    FFR_VAR_GAUSSIAN_BLUR,   /* GaussianBlur :352  -                                   */
    FFR_VAR_SQUARE_NOISE,    /* SquareNoise :369   -                                   */
    FFR_VAR_SEPARATION,      /* Separation :383    [i]=sep^2 [4+i]=inside              */
    FFR_VAR_SPLITS,          /* Splits :410        [i]=params                          */
    FFR_VAR_PRE_BLUR,        /* PreBlur :434       -                                   */
    FFR_VAR_MODULUS,         /* Modulus :451       [i]=2*param [4+i]=1/(2*param)       */
    FFR_VAR_CELLN,           /* CellN :476         [i]=sizes [4+i]=1/sizes             */
    FFR_VAR_SWIRL,           /* Swirl :510         -                                   */
    FFR_VAR_HORSESHOE,       /* Horseshoe :528     -                                   */
    FFR_VAR_POLAR,           /* Polar :546         -                                   */
    FFR_VAR_POLAR2,          /* Polar2 :561        -                                   */
    FFR_VAR_HANDKERCHIEF,    /* Handkerchief :575  -                                   */
    FFR_VAR_HEART,           /* Heart :590         -                                   */
    FFR_VAR_DISC,            /* Disc :607          -                                   */
    FFR_VAR_DISC2,           /* Disc2 :624         rotpi, addval.x, addval.y           */
    FFR_VAR_WAVES,           /* Waves :660         xf, xs, yf, ys                      */
    FFR_VAR_FAN,             /* Fan :685           dx, dy                              */
    FFR_VAR_RINGS,           /* Rings :714         dx                                  */
    FFR_VAR_SPIRAL,          /* Spiral :738        -                                   */
    FFR_VAR_HYPERBOLIC,      /* Hyperbolic :757    -                                   */
    FFR_VAR_DIAMOND,         /* Diamond :772       -                                   */
    FFR_VAR_EX,              /* Ex :789            -                                   */
    FFR_VAR_JULIA,           /* Julia :808         -                                   */
    FFR_VAR_EXPONENTIAL,     /* Exponential :825   -                                   */
    FFR_VAR_POWER,           /* Power :843         -                                   */
    FFR_VAR_COSINE,          /* Cosine :858        -                                   */
    FFR_VAR_BLOB,            /* Blob :875          mid, amp, waves                     */
    FFR_VAR_PDJ,             /* PDJ :901           a, b, c, d                          */
    FFR_VAR_CYLINDER,        /* Cylinder :930      -                                   */
    FFR_VAR_PERSPECTIVE,     /* Perspective :943   dist, vsin, vfcos                   */
    FFR_VAR_JULIAN,          /* JuliaN :967        abspower, invpower, cn              */
    FFR_VAR_JULIASCOPE,      /* JuliaScope :994    abspower, invpower, cn              */
    FFR_VAR_RADIAL_BLUR,     /* RadialBlur :1023   spin, zoom, flam3weight             */
    FFR_VAR_PIE,             /* Pie :1050          slices, rotation, thickness, invslices2pi */
    FFR_VAR_NGON,            /* NGon :1077         powerval, angle, corners, circle, invangle */
    FFR_VAR_CURL,            /* Curl :1107         c1, c2                              */
    FFR_VAR_ARCH,            /* Arch :1134         flam3weight                         */
    FFR_VAR_TANGENT,         /* Tangent :1156      -                                   */
    FFR_VAR_RAYS,            /* Rays :1172         flam3weight                         */
    FFR_VAR_BLADE,           /* Blade :1196        flam3weight                         */
    FFR_VAR_SECANT,          /* Secant :1218       flam3weight                         */
    FFR_VAR_TWINTRIAN,       /* Twintrian :1240    flam3weight                         */
    FFR_VAR_CROSS,           /* Cross :1265        -                                   */
    FFR_VAR_EXP,             /* Exp :1282          -                                   */
    FFR_VAR_LOG,             /* Log :1300          -                                   */
    FFR_VAR_SIN,             /* Sin :1313          -                                   */
    FFR_VAR_COS,             /* Cos :1332          -                                   */
    FFR_VAR_TAN,             /* Tan :1351          -                                   */
    FFR_VAR_SEC,             /* Sec :1371          -                                   */
    FFR_VAR_CSC,             /* Csc :1391          -                                   */
    FFR_VAR_COT,             /* Cot :1411          -                                   */
    FFR_VAR_SINH,            /* Sinh :1431         -                                   */
    FFR_VAR_COSH,            /* Cosh :1450         -                                   */
    FFR_VAR_TANH,            /* Tanh :1469         -                                   */
    FFR_VAR_SECH,            /* Sech :1489         -                                   */
    FFR_VAR_CSCH,            /* Csch :1509         -                                   */
    FFR_VAR_COTH,            /* Coth :1529         -                                   */
    FFR_VAR_AUGER,           /* Auger :1549        freq, augerweight, scale/2, sym     */
    FFR_VAR_FLUX,            /* Flux :1576         2+spread, fluxweight                */
    FFR_VAR_MOBIUS,          /* Mobius :1605       a.x a.y b.x b.y c.x c.y d.x d.y     */
    FFR_VAR_SCRY,            /* Scry :1635         scryweight                          */
    FFR_VAR_SPLIT,           /* Split :1655        xsize*pi, ysize*pi                  */
    FFR_VAR_STRIPES,         /* Stripes :1678      1-space, warp                       */
    FFR_VAR_WEDGE,           /* Wedge :1701        swirl, count, angle, hole, cf       */
    FFR_VAR_WEDGE_JULIA,     /* WedgeJulia :1729   cn, abspower, invpower, count, angle, cf */
    FFR_VAR_WEDGE_SPH,       /* WedgeSph :1761     swirl, count, cf, angle, hole       */
    FFR_VAR_WHORL,           /* Whorl :1790        inside, outside, whorlweight        */
    FFR_VAR_SUPERSHAPE,      /* Supershape :1815   pm_4, pneg1_n1, n2, n3, rnd, holes  */
    FFR_VAR_FLOWER,          /* Flower :1847       petals, holes                       */
    FFR_VAR_CONIC,           /* Conic :1869        eccen, holes                        */
    FFR_VAR_PARABOLA,        /* Parabola :1891     h, w                                */
    FFR_VAR_BIPOLAR,         /* Bipolar :1914      -pi/2*shift                         */
    FFR_VAR_BOARDERS,        /* Boarders :1940     prob                                */
    FFR_VAR_BUTTERFLY,       /* Butterfly :1988    -                                   */
    FFR_VAR_CELL,            /* Cell :2008         size, 1/size                        */
    FFR_VAR_CPOW,            /* CPow :2038         va, vc, vd, power                   */
    FFR_VAR_CURVE,           /* Curve :2066        invxl, invyl, xamp, yamp            */
    FFR_VAR_EDISC,           /* EDisc :2096        -                                   */
    FFR_VAR_ELLIPTIC,        /* Elliptic :2125     -                                   */
    FFR_VAR_ESCHER,          /* Escher :2149       vc, vd                              */
    FFR_VAR_FOCI,            /* Foci :2176         -                                   */
    FFR_VAR_LAZYSUSAN,       /* LazySusan :2197    px, py, spin, twist, space, lsweight */
    FFR_VAR_LOONIE,          /* Loonie :2235       loonieweight, loonieweight^2        */
    FFR_VAR_OSCOPE,          /* OScope :2259       2*pi*freq, amp, damp, sep           */
    FFR_VAR_POPCORN,         /* Popcorn :2286      px, py, pc                          */
    FFR_VAR_SPHERICAL_P,     /* SphericalP :2312   norm                                */
    FFR_VAR_UNIT_SPHERE,     /* UnitSphere :2333   -                                   */
    FFR_VAR_UNIT_SPHERE_P,   /* UnitSphereP :2347  norm                                */
    FFR_VAR_UNIT_CUBE,       /* UnitCube :2368     -                                   */
    FFR_VAR_COUNT_PLUS_1
};
#define FFR_VAR_COUNT (FFR_VAR_COUNT_PLUS_1 - 1)          /* 98 */
#define FFR_VAR_FIRST_2D FFR_VAR_SWIRL                    /* 17 */
#define FFR_VAR_LAST_2D  FFR_VAR_POPCORN                  /* 94 */

/* One variation of an xform, after the reference constructor ran
   (Variation :38-59, VariationFrom2D :63-107). */
typedef struct ffr_variation
{
    uint32_t op;         /* enum ffr_var_op */
    uint32_t axis_x;     /* 2-D variations lifted to dims>2 (variations.hpp:98-105) */
    uint32_t axis_y;
    uint32_t reserved;
    double weight;       /* Variation::getWeight() */
    double params[FFR_MAX_VAR_PARAMS];
} ffr_variation;

/* One xform after XForm::XForm + _optimize (types/xform.hpp:47-172): zero-weight
   variations removed, JSON order kept. Matrices are row-major dims x dims in the
   leading entries (A[i*dims+j]); absent affines hold the identity like the
   reference's default Affine (affine.hpp:31-36). */
typedef struct ffr_xform
{
    uint64_t id;               /* index in the JSON "xforms" array; FFR_FINAL_XFORM_ID for the final xform */
    double weight;             /* 1.0 for the final xform (xform.hpp:88) */
    uint32_t has_pre;
    uint32_t has_post;
    uint32_t has_color;
    uint32_t num_vars;
    double pre_A[9], pre_b[3];
    double post_A[9], post_b[3];
    double color_speed;
    const double *color;       /* color_dims entries when has_color, else NULL */
    const ffr_variation *vars; /* num_vars entries */
} ffr_xform;

/* POD flatten of Flame<dims> after _optimize + _setupCumulativeWeights
   (types/flame.hpp:44-83): zero-weight xforms dropped, the rest sorted by
   decreasing weight with the same std::sort, cumulative weights computed with the
   same recurrence, last entry forced to 1.0. The caller owns it; ffr_cuda_create
   copies what it needs. */
typedef struct ffr_flame_desc
{
    uint32_t dims;             /* 1..3 */
    uint32_t color_dims;       /* r, 0..127 */
    uint32_t elem_size;        /* sizeof(num_t)==sizeof(hist_t), types.hpp:24-41: 8 = double/uint64_t
                                  (as shipped), 4 = float/uint32_t; every value of the desc must
                                  then be exactly representable in that num_t */
    uint32_t has_final;
    uint64_t size[FFR_MAX_DIMS];
    double bounds_lo[FFR_MAX_DIMS];
    double bounds_hi[FFR_MAX_DIMS];
    uint32_t num_xforms;       /* after optimisation */
    uint32_t num_xform_ids;    /* Flame::getXFormIDCount(): xforms in the JSON */
    const ffr_xform *xforms;   /* num_xforms entries, selection order */
    const double *xfcw;        /* num_xforms cumulative weights */
    const ffr_xform *final_xform; /* NULL unless has_final */
} ffr_flame_desc;

/* Mirror of BufferRenderer::stats (buffer_renderer.hpp:85-106). Counters are
   cumulative over the life of the context like the reference's; the bad value
   lists are those of the last render call (render() clears them, :277-278). */
typedef struct ffr_stats
{
    uint64_t s_iter;                       /* samples iterated */
    uint64_t s_plot;                       /* samples plotted  */
    uint64_t xf_dist[FFR_MAX_XFORMS];      /* by JSON xform id */
    double pt_min[FFR_MAX_DIMS];           /* extremes of p (not pf), :188-194 */
    double pt_max[FFR_MAX_DIMS];
    uint64_t n_bad;                        /* bad values hit in the last call */
    uint64_t bad_xf[FFR_MAX_BAD_RECORDED]; /* first FFR_MAX_BAD_RECORDED of them; */
    double bad_pt[FFR_MAX_BAD_RECORDED][FFR_MAX_DIMS]; /* order across chains is unspecified */
} ffr_stats;

/* Scatter strategies of the accumulate step. */
enum ffr_scatter_mode
{
    FFR_SCATTER_AUTO = 0,
    FFR_SCATTER_GLOBAL = 1,     /* one RED per element straight to L2/HBM */
    FFR_SCATTER_WARP_AGG = 2,   /* __match_any_sync aggregation of colliding lanes first */
    FFR_SCATTER_SMEM_TILE = 3,  /* reserved, NOT implemented: ffr_cuda_create_ex fails with a message.
                                   A shared-memory privatised tile has no room next to the chains'
                                   ISAAC state (scatter strategies); what the library does
                                   instead for small/mid buffers is the L2-resident accumulation tile
                                   of the pure-affine kernel, chosen automatically */
    FFR_SCATTER_TRACE = 5,      /* internal to ffr_cuda_atomic_roofline patterns 1 and 2 */
    FFR_SCATTER_DISCARD = 4     /* diagnostic: iterate and count but issue no REDs (measures the
                                   compute-only rate for the roofline analysis; buffer untouched) */
};

typedef struct ffr_options
{
    uint32_t struct_size;       /* sizeof(ffr_options), for ABI growth */
    uint32_t scatter_mode;      /* enum ffr_scatter_mode */
    uint32_t regroup;           /* 0 auto, 1 force off, 2 force on: per-iteration xform regrouping */
    uint32_t blocks_per_sm;     /* 0 auto */
    void *external_buffer;      /* device pointer (ndev must be 1): render into caller-owned
                                   memory (e.g. a torch tensor) instead of allocating */
    void *stream;               /* cudaStream_t to launch on (ndev must be 1); NULL = own stream */
    uint32_t jit;               /* flame-specialised kernel compiled at run time (NVRTC): 0 auto (lazily,
                                   by the first render call large enough to repay the compile), 1 never,
                                   2 at create (fails if unavailable) */
    uint32_t reserved0;
} ffr_options;

typedef struct ffr_ctx ffr_ctx;

/* progress callback: called from the calling thread only, done/total in chains */
typedef void (*ffr_progress_cb)(void *user, uint64_t chains_done, uint64_t chains_total);

/* return codes */
#define FFR_OK            0
#define FFR_BAD_VALUES    1   /* bad value limit exceeded: render() returned false */
#define FFR_E_INVALID    -1   /* invalid argument (reference: std::runtime_error) */
#define FFR_E_CUDA       -2   /* CUDA runtime error */
#define FFR_E_NODEVICE   -3   /* no usable sm_100 device; there is NO CPU fallback */
#define FFR_E_UNSUPPORTED -4

const char *ffr_cuda_version(void);
int ffr_cuda_device_count(void);

/* SplitMix64 of (base_seed + k): the u64 handed to Isaac::setSeed(u64)
   (rng/isaac.hpp:267-271) for chain k. */
uint64_t ffr_chain_seed(uint64_t base_seed, uint64_t chain_index);

/* BufferRenderer(const Flame&) (buffer_renderer.hpp:254, _init :114-140). devices ==
   NULL means devices 0..ndev-1. Returns NULL and fills err on failure ("histogram too
   big" for cells >= 2^48 like :132-133). */
ffr_ctx *ffr_cuda_create(const ffr_flame_desc *desc, const int *devices, int ndev,
        char *err, size_t errlen);
ffr_ctx *ffr_cuda_create_ex(const ffr_flame_desc *desc, const int *devices, int ndev,
        const ffr_options *opt, char *err, size_t errlen);
void ffr_cuda_destroy(ffr_ctx *ctx);
const char *ffr_cuda_last_error(const ffr_ctx *ctx);

/* bytes of the buffer: cells * (1 + color_dims) * elem_size (buffer_renderer.hpp:137) */
size_t ffr_cuda_buffer_bytes(const ffr_ctx *ctx);
uint64_t ffr_cuda_buffer_cells(const ffr_ctx *ctx);
/* device pointer of device #dev_index's private buffer (reference layout) */
void *ffr_cuda_device_buffer(ffr_ctx *ctx, int dev_index);

/* addBuffer (buffer_renderer.hpp:375-452): counts add as u64, colours as f64.
   bytes must equal ffr_cuda_buffer_bytes. */
int ffr_cuda_add_buffer(ffr_ctx *ctx, const void *host, size_t bytes);
int ffr_cuda_clear_buffer(ffr_ctx *ctx);

/* render (buffer_renderer.hpp:269-338): samples split into ceil(samples/chain_len)
   chains sharded contiguously over the context's devices. chain_len >= 256 like the
   reference's batch_size check (:283-285). Blocking. Returns FFR_OK, FFR_BAD_VALUES
   or < 0. stats may be NULL. */
int ffr_cuda_render(ffr_ctx *ctx, uint64_t samples, uint64_t chain_len, uint64_t base_seed,
        uint64_t bv_limit, ffr_progress_cb cb, void *user, ffr_stats *stats);

/* The sharding primitive: render chains [chain_first, chain_first+chain_count) of a job
   whose chains have chain_len samples, the last chain of THIS range having last_len
   (0 = chain_len). Used by one-process-per-GPU hosts (bench.py under torchrun). */
int ffr_cuda_render_chains(ffr_ctx *ctx, uint64_t chain_first, uint64_t chain_count,
        uint64_t chain_len, uint64_t last_len, uint64_t base_seed, uint64_t bv_limit,
        ffr_stats *stats);

/* Same, but only enqueues the kernels on the context's stream and returns; pair with
   ffr_cuda_sync / ffr_cuda_get_stats. Lets a caller time the launches with events on
   its own stream (ffr_options.stream). Single device contexts only. */
int ffr_cuda_render_chains_async(ffr_ctx *ctx, uint64_t chain_first, uint64_t chain_count,
        uint64_t chain_len, uint64_t last_len, uint64_t base_seed, uint64_t bv_limit);
/* The rest of the streaming interface (single device contexts): clear, -i add and read-back that
   only enqueue on the context's stream, so that a host running two contexts on two streams
   overlaps one context's PCIe legs with the other's render (bench.py's e2e loop does). The host
   buffers must be PAGE-LOCKED (cudaHostAlloc / cudaHostRegister / torch pin_memory) and stay
   valid and untouched until ffr_cuda_sync returns; pageable memory is refused (FFR_E_INVALID).
   The upload of the add runs on a copy engine (a stream of its own, ordered by events) into a
   full-size staging buffer kept by the context. */
int ffr_cuda_clear_buffer_async(ffr_ctx *ctx);
int ffr_cuda_add_buffer_async(ffr_ctx *ctx, const void *pinned_host, size_t bytes);
int ffr_cuda_read_buffer_async(ffr_ctx *ctx, void *pinned_host, size_t bytes);
int ffr_cuda_sync(ffr_ctx *ctx);
int ffr_cuda_get_stats(ffr_ctx *ctx, ffr_stats *stats);
/* chains the render kernel keeps in flight per device (SMs x resident blocks x 256): size
   chain counts as a multiple of this for full waves */
uint64_t ffr_cuda_resident_chains(const ffr_ctx *ctx);
/* number of kernel launches issued by this context so far */
uint64_t ffr_cuda_launch_count(const ffr_ctx *ctx);

/* ---- the flame-specialised render kernel (run-time compiled; csrc/ffr_jit_kernel.cuh) ----
   The reference dispatches every variation through a virtual call (variations.hpp:38-59,
   xform.hpp:218-221); the ahead-of-time kernels interpret a flattened op list; this path turns
   the flame into straight-line sm_100a code with NVRTC. Results are identical to the
   interpreter kernels bit for bit (same device functions, same order, -fmad=false). */
typedef struct ffr_jit_info
{
    uint32_t active;            /* renders of this context run the compiled kernel */
    uint32_t eligible, failed, from_cache;
    uint32_t threads_per_block, slots_per_block, blocks_per_sm, registers;
    uint64_t smem_bytes, cubin_bytes, source_bytes;
    double compile_seconds;     /* NVRTC time of this context's compile (0 when cached) */
    char message[256];          /* why it is unavailable, if it is */
} ffr_jit_info;
int ffr_cuda_jit_info(const ffr_ctx *ctx, ffr_jit_info *info);
/* compile + load now (what ffr_options.jit = 2 does at create) */
int ffr_cuda_jit_enable(ffr_ctx *ctx);
/* the generated CUDA source (empty until compiled); returns its length */
size_t ffr_cuda_jit_source(const ffr_ctx *ctx, char *buf, size_t buflen);
/* Generate and compile for a flame WITHOUT a device (NVRTC needs none): the build check of
   the run-time path. source may be NULL. */
int ffr_cuda_jit_compile(const ffr_flame_desc *desc, char *source, size_t source_len,
        size_t *cubin_bytes, char *err, size_t errlen);

/* Sum the per-device private buffers into device 0's buffer over NVLink peer memory
   (no-op for one device): a reduce-scatter -- every device adds its 1/N slice of all peers'
   buffers, reading them in place, all devices at once -- then device 0 gathers the finished
   slices. ffr_cuda_read_buffer calls it when needed. */
int ffr_cuda_reduce(ffr_ctx *ctx);

/* The same typed slice sum for hosts that run one process per GPU and exchange buffer slices
   themselves (an all-to-all over NCCL): dst[i] += sum_k srcs[k][i] for n_elems elements, all
   DEVICE pointers on the context's device, element i being element first_elem + i of the buffer
   (element 0 of each cell adds as a count, the others as colour sums). n_src <= 15. Launched on
   the context's stream, not synchronised. */
int ffr_cuda_sum_device_slices(ffr_ctx *ctx, void *dst, const void *const *srcs, int n_src,
        uint64_t first_elem, uint64_t n_elems);

/* One process per GPU on one box (what `ffr-buf.out --gpus N` does: every CUDA context is created
   by a process of its own, in parallel, instead of N contexts serialised inside one process):
   a worker exports its single-device context's buffer as a CUDA IPC handle (64 bytes, sent to the
   collecting process through a pipe), the collector adds the workers' buffers to its own buffer
   by reading them in place over NVLink peer memory (K2d, typed by cell position). The worker must
   keep its context alive until the collector's call returned. n <= 15. */
#define FFR_IPC_HANDLE_BYTES 64
int ffr_cuda_ipc_export(ffr_ctx *ctx, void *handle);
int ffr_cuda_ipc_add(ffr_ctx *ctx, const void *handles, int n);

/* writeBuffer (buffer_renderer.hpp:476-480): the reduced buffer in the reference file
   layout: cells x [count u64, c0..c(r-1) f64], dimension 0 fastest, native endian. */
int ffr_cuda_read_buffer(ffr_ctx *ctx, void *host, size_t bytes);

/* histogramSum / histogramMax (buffer_renderer.hpp:483-509) computed on the device */
int ffr_cuda_histogram_sum_max(ffr_ctx *ctx, uint64_t *sum, uint64_t *max);

/* ---- log-density tone map: the pixel math of ffr-img (src/ffr_img.cpp:199-309,
   renderers/image_renderer.hpp:112-192), 2-d flames only ---- */
enum ffr_tonemap_mode
{
    FFR_TONE_MONO = 1,   /* -m: count != 0 ? 1 : 0, 8-bit gray            ffr_img.cpp:259-265 */
    FFR_TONE_GRAY = 2,   /* -g: pow(log(1+n)/max, 1/gamma)                :236-243,267-279 */
    FFR_TONE_RGB  = 3    /* -c: that times colour_i/n, 3 colour dims only :280-305 */
};

typedef struct ffr_tonemap_info
{
    uint64_t hist_min, hist_max;     /* "histogram bounds" line, ffr_img.cpp:217-218 */
    double scaler_min, scaler_max;   /* "scaler bounds" line: log(1+n) extremes, :229-230 */
    uint32_t width, height, channels, bits;
} ffr_tonemap_info;

/* Sums the device buffers (like ffr_cuda_read_buffer), reduces min/max, maps every cell to a
   pixel on the device and copies the image to `pixels`: height rows of width pixels, row y =
   buffer dimension 1, `channels` samples per pixel, samples of `bits` (8 or 16) bits in HOST
   byte order. bytes must be width*height*channels*bits/8. Errors follow ffr-img: "gamma too
   small" (< 1e-20), "bits per channel must be 8 or 16", "buffer must use 3 color dimensions",
   "only 2D flames supported", "histogram is (probably) empty". */
int ffr_cuda_tonemap(ffr_ctx *ctx, int mode, int bits, double gamma, void *pixels, size_t bytes,
        ffr_tonemap_info *info);

/* Test hook on the same device code as the render kernels: for each of n points, seed
   an ISAAC stream with seeds[i] (Isaac::setSeed(u64)) and apply xform #xf_index (sorted
   order; -1 = final xform) once: XForm::applyIteration (types/xform.hpp:211-227).
   Host pointers, n*dims doubles in and out. */
int ffr_cuda_iterate_points(ffr_ctx *ctx, int64_t xf_index, uint64_t n, const uint64_t *seeds,
        const double *pts_in, double *pts_out);
/* Test hook: first n words of the ISAAC-64 stream after setSeed(seed) (isaac.hpp:267-329) */
int ffr_cuda_isaac_words(ffr_ctx *ctx, uint64_t seed, uint64_t n, uint64_t *out);

/* Atomic-scatter microbenchmarks: the measured scatter roofline. Bare REDs
   (1 count + color_dims colour sums per cell) issued from 2048 resident threads per SM with
   nothing to wait for; returns elapsed ms (CUDA events) and in *n_done the cells actually hit.
   pattern 0 = n_atomics uniformly random cells of the context's buffer (a reference point for
     the memory system at this buffer size, not a ceiling for a render);
   pattern 1 = replay of the flame's own scatter: a render of resident_chains chains x
     (n_atomics / resident_chains) samples records the address every plotted sample was scattered
     to (the buffer cell, or the cell of the pure-affine kernel's own accumulation tile after its
     scramble / row directory), then the recorded stream is replayed, streamed from memory;
   pattern 2 = the same replay from shared-memory windows of the trace, each replayed many times,
     so that the timed loop issues nothing but REDs (the streamed trace costs an L2-resident
     scatter a fifth of the L2's sector rate).
   The better of patterns 1 and 2 is the scatter ceiling of that render. Buffer contents are
   garbage afterwards; statistics are restored. */
int ffr_cuda_atomic_roofline(ffr_ctx *ctx, uint64_t n_atomics, int pattern, float *ms);
int ffr_cuda_atomic_roofline_ex(ffr_ctx *ctx, uint64_t n_atomics, int pattern, float *ms,
        uint64_t *n_done);

#ifdef __cplusplus
}
#endif

#endif /* FFR_CUDA_H */
