/*
ffr_buffer_renderer.hpp -- the reference's BufferRenderer<dims> surface over libffr_cuda
(header only, C++17; link with -lffr_cuda).

A C++ host that used tkoz::flame::BufferRenderer (src/renderers/buffer_renderer.hpp:40-585)
switches to the B200 path by constructing this class instead: same method names, same
argument meaning, same return values, same std::runtime_error texts. Everything forwards to
the C ABI (ffr_cuda.h, ffr_flame.h); there is no CPU path behind it -- without a usable device
the constructor throws.

What is the same
  render(num_samples, num_threads, batch_size, bv_limit, cb_batch, cb_thread)   :269-338
      throws on threads == 0, threads > 65535, batch_size < 256 (same messages); false when the
      bad value limit was exceeded; cb_batch is called once per finished batch (= chain) from
      the calling thread; cb_thread once per worker, before the first batch.
  renderSeeded(num_samples, batch_size, bv_limit, cb_batch)                    :349-375
  addBuffer(ptr | vector | renderer | istream), readBuffer(istream), writeBuffer(ostream)
                                                                                :377-480
  histogramSum / Min / Max, getBuffer*, getSamplesIterated / Plotted, getXFormDistribution,
  getBadValueXForms / Points, getPointExtremes, getDimMults, getIndexMults, getDims,
  getColorDims                                                                  :482-583
  buffer layout: cells x [count, c0..c(r-1)], dimension 0 fastest, native endian.

What differs, and why
  * construction takes the flame JSON TEXT (the Flame<dims> class lives behind ffr_flame.h);
    parse and validation errors are JsonError with the reference's messages.
  * seeding. The reference seeds each worker thread from the clock (:300) and hands batches
    out dynamically, so its output is not reproducible; here batch k of a call is always the
    chain seeded splitmix64(seed + k) -- a pure function of (seed, batch_size, num_samples),
    whatever the number of devices. setSeed() chooses the seed (default: from the clock, like
    the reference); every render call then advances it by the number of batches it used, so
    consecutive calls never repeat a chain. renderSeeded is the same render (the reference's
    continues the calling thread's generator across batches instead, :358-371; both are "one
    deterministic stream per call").
  * workers are GPUs: num_threads is validated like the reference's and otherwise unused;
    cb_thread receives a default-constructed std::thread (there is no host thread to name) and
    the device's index.
  * enable_color = false is not provided (ffr_buf.cpp never instantiates it).
  * the buffer lives on the device: getBuffer()/getBufferCell()/histogramMin() fetch a host copy
    on first use after a change.
*/

#ifndef FFR_BUFFER_RENDERER_HPP
#define FFR_BUFFER_RENDERER_HPP

#include "ffr_cuda.h"
#include "ffr_flame.h"

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <istream>
#include <memory>
#include <ostream>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

namespace ffr
{

/* utils/json.hpp:13-25 */
struct JsonError : public std::runtime_error
{
    explicit JsonError(const std::string &what): std::runtime_error(what) {}
};

typedef double num_t;       /* types/types.hpp:24-41, the build as shipped */
typedef uint64_t hist_t;

/* buffer_renderer.hpp:20-29 without the atomic members (the atomics are the device's) */
union buf_elem_t
{
    hist_t uintval;
    num_t floatval;
    buf_elem_t(): uintval(0) {}
};
static_assert(sizeof(buf_elem_t) == sizeof(hist_t),"buffer element is one word");

struct DeviceOptions
{
    std::vector<int> devices;   /* CUDA ordinals; empty = the first `num_devices` */
    int num_devices = 1;
    int jit = 0;                /* ffr_options.jit: 0 auto, 1 never, 2 at construction */
};

template <size_t dims, bool enable_color = true>
class BufferRenderer
{
    static_assert(dims > 0 && dims <= FFR_MAX_DIMS,"ffr_buf.cpp:134-142 instantiates dims 1, 2, 3");
    static_assert(enable_color,"the device renderer always keeps the flame's colour dimensions");

public:
    typedef std::array<num_t,dims> point_t;
    typedef std::pair<num_t,num_t> num_pair_t;

private:
    struct FlameDel { void operator()(ffr_flame *f) const { ffr_flame_free(f); } };
    struct CtxDel { void operator()(ffr_ctx *c) const { ffr_cuda_destroy(c); } };
    std::unique_ptr<ffr_flame,FlameDel> flame;
    std::unique_ptr<ffr_ctx,CtxDel> ctx;
    const ffr_flame_desc *desc = nullptr;
    size_t buffer_cells = 0, buffer_cell_size = 0, ndev = 1;
    uint64_t seed = 0;
    std::array<num_t,dims> mult_d{};
    std::array<size_t,dims> mult_i{};

    struct
    {
        size_t s_iter = 0, s_plot = 0;
        std::vector<size_t> xf_dist, bv_xfs;
        std::vector<point_t> bv_pts;
        std::array<num_pair_t,dims> pt_max{};
    }
    stats;
    std::unique_ptr<ffr_stats> raw{new ffr_stats};

    mutable std::vector<buf_elem_t> host;   /* copy of the device buffer, fetched on demand */
    mutable bool host_valid = false;

    [[noreturn]] void fail(const char *where) const
    {
        throw std::runtime_error(std::string(where) + ": " + ffr_cuda_last_error(ctx.get()));
    }

    void take_stats()
    {
        stats.s_iter = raw->s_iter;
        stats.s_plot = raw->s_plot;
        stats.xf_dist.assign(raw->xf_dist,raw->xf_dist + desc->num_xform_ids);
        const size_t nb = (size_t)std::min<uint64_t>(raw->n_bad,FFR_MAX_BAD_RECORDED);
        stats.bv_xfs.assign(raw->bad_xf,raw->bad_xf + nb);
        stats.bv_pts.resize(nb);
        for (size_t i = 0; i < nb; ++i)
            for (size_t d = 0; d < dims; ++d)
                stats.bv_pts[i][d] = raw->bad_pt[i][d];
        for (size_t d = 0; d < dims; ++d)
            stats.pt_max[d] = num_pair_t(raw->pt_min[d],raw->pt_max[d]);
    }

    struct Progress
    {
        const std::function<void()> *cb;
        uint64_t reported;
    };
    static void on_progress(void *user, uint64_t done, uint64_t)
    {
        Progress *p = (Progress*)user;
        for (; p->reported < done; ++p->reported)
            (*p->cb)();
    }

    bool run(size_t num_samples, size_t batch_size, size_t bv_limit, const std::function<void()> &cb_batch)
    {
        Progress pr{&cb_batch,0};
        const int rc = ffr_cuda_render(ctx.get(),num_samples,batch_size,seed,bv_limit,
            cb_batch ? &on_progress : nullptr,&pr,raw.get());
        host_valid = false;
        if (rc < 0)
            fail("BufferRenderer::render()");
        seed += (num_samples + batch_size - 1) / batch_size;
        take_stats();
        return rc == FFR_OK;
    }

    const std::vector<buf_elem_t> &fetch() const
    {
        if (!host_valid)
        {
            host.resize(buffer_cells*buffer_cell_size);
            if (ffr_cuda_read_buffer(ctx.get(),host.data(),host.size()*sizeof(buf_elem_t)) != FFR_OK)
                fail("BufferRenderer::getBuffer()");
            host_valid = true;
        }
        return host;
    }

public:
    /* Flame<dims>(json) + BufferRenderer(flame), ffr_buf.cpp:149-150 */
    explicit BufferRenderer(const std::string &flame_json, const DeviceOptions &opt = DeviceOptions())
    {
        char err[1024] = "";
        flame.reset(ffr_flame_from_json(flame_json.data(),flame_json.size(),err,sizeof err));
        if (!flame)
            throw JsonError(err);
        desc = ffr_flame_get_desc(flame.get());
        if (desc->dims != dims)
            throw std::runtime_error("BufferRenderer: the flame has " + std::to_string(desc->dims) +
                " dimensions, the renderer " + std::to_string(dims));
        double md[FFR_MAX_DIMS];
        uint64_t mi[FFR_MAX_DIMS], cells = 0, cell = 0;
        if (ffr_flame_layout(desc,md,mi,&cells,&cell) != FFR_OK)
            throw std::runtime_error("BufferRenderer(): histogram too big");       /* :132-133 */
        for (size_t d = 0; d < dims; ++d)
        {
            mult_d[d] = md[d];
            mult_i[d] = (size_t)mi[d];
        }
        buffer_cells = (size_t)cells;
        buffer_cell_size = (size_t)cell;
        ffr_options o;
        std::memset(&o,0,sizeof o);
        o.struct_size = sizeof o;
        o.jit = (uint32_t)opt.jit;
        ndev = opt.devices.empty() ? (size_t)std::max(1,opt.num_devices) : opt.devices.size();
        ctx.reset(ffr_cuda_create_ex(desc,opt.devices.empty() ? nullptr : opt.devices.data(),(int)ndev,&o,
            err,sizeof err));
        if (!ctx)
            throw std::runtime_error(err);
        std::memset(raw.get(),0,sizeof(ffr_stats));
        take_stats();
        for (size_t d = 0; d < dims; ++d)       /* :134-135: before the first sample */
            stats.pt_max[d] = num_pair_t(INFINITY,-INFINITY);
        seed = (uint64_t)std::chrono::high_resolution_clock::now().time_since_epoch().count();
    }

    /* the seed of the NEXT render call's first batch (see the header comment) */
    void setSeed(uint64_t s) { seed = s; }
    uint64_t getSeed() const { return seed; }

    bool render(size_t num_samples, size_t num_threads, size_t batch_size, size_t bv_limit,
            std::function<void()> cb_batch = nullptr,
            std::function<void(const std::thread&,size_t)> cb_thread = nullptr)
    {
        stats.bv_pts.clear();
        stats.bv_xfs.clear();
        if (num_samples == 0)
            return true;
        if (num_threads == 0)
            throw std::runtime_error("BufferRenderer::render(): threads must be positive");
        if (num_threads > 65535)
            throw std::runtime_error("BufferRenderer::render(): too many threads");
        if (batch_size < 256)
            throw std::runtime_error("BufferRenderer::render(): batch size too small");
        if (cb_thread)
        {
            const std::thread none;
            for (size_t i = 0; i < ndev; ++i)
                cb_thread(none,i);
        }
        return run(num_samples,batch_size,bv_limit,cb_batch);
    }

    bool renderSeeded(size_t num_samples, size_t batch_size, size_t bv_limit,
            std::function<void()> cb_batch = nullptr)
    {
        stats.bv_pts.clear();
        stats.bv_xfs.clear();
        if (num_samples == 0)
            return true;
        if (batch_size == 0)
            throw std::runtime_error("BufferRenderer::render(): batch size must be positive");
        if (batch_size < 256)   /* the device kernels' chains are the reference's render() batches */
            throw std::runtime_error("BufferRenderer::render(): batch size too small");
        return run(num_samples,batch_size,bv_limit,cb_batch);
    }

    template <typename T>
    void addBuffer(const T *buf)
    {
        static_assert(sizeof(T) == sizeof(buf_elem_t),"buffer elements are one word");
        if (ffr_cuda_add_buffer(ctx.get(),buf,buffer_cells*buffer_cell_size*sizeof(buf_elem_t)) != FFR_OK)
            fail("BufferRenderer::addBuffer()");
        host_valid = false;
    }

    template <typename T>
    void addBuffer(const std::vector<T> &buf)
    {
        static_assert(sizeof(T) == sizeof(buf_elem_t),"buffer elements are one word");
        if (buf.size() != buffer_cells*buffer_cell_size)
            throw std::runtime_error("BufferRenderer::addBuffer(): sizes do not match");
        addBuffer(buf.data());
    }

    void addBuffer(const BufferRenderer<dims,enable_color> &renderer)
    {
        if (buffer_cells != renderer.buffer_cells || buffer_cell_size != renderer.buffer_cell_size)
            throw std::runtime_error("BufferRenderer::addBuffer(): formats do not match");
        addBuffer(renderer.getBuffer());
    }

    /* false on a short read. The reference's default (alloc_tmp = false) may leave a partial
       sum behind in that case; here the buffer is untouched either way. */
    template <bool alloc_tmp = false>
    bool addBuffer(std::istream &is)
    {
        std::vector<buf_elem_t> buf(buffer_cells*buffer_cell_size);
        is.read((char*)buf.data(),(std::streamsize)(buf.size()*sizeof(hist_t)));
        if (!is.good())
            return false;
        addBuffer(buf);
        return true;
    }

    template <bool alloc_tmp = false>
    bool readBuffer(std::istream &is)
    {
        std::vector<buf_elem_t> buf(buffer_cells*buffer_cell_size);
        is.read((char*)buf.data(),(std::streamsize)(buf.size()*sizeof(hist_t)));
        if (!is.good())
            return false;
        if (ffr_cuda_clear_buffer(ctx.get()) != FFR_OK)
            fail("BufferRenderer::readBuffer()");
        addBuffer(buf);
        return true;
    }

    bool writeBuffer(std::ostream &os) const
    {
        const std::vector<buf_elem_t> &b = fetch();
        os.write((const char*)b.data(),(std::streamsize)(b.size()*sizeof(hist_t)));
        return os.good();
    }

    [[nodiscard]] size_t histogramSum() const
    {
        uint64_t sum = 0, mx = 0;
        if (ffr_cuda_histogram_sum_max(ctx.get(),&sum,&mx) != FFR_OK)
            fail("BufferRenderer::histogramSum()");
        return (size_t)sum;
    }

    [[nodiscard]] hist_t histogramMin() const
    {
        const std::vector<buf_elem_t> &b = fetch();
        hist_t ret = (hist_t)-1;
        for (size_t i = 0; i < buffer_cells; ++i)
            ret = std::min(ret,b[i*buffer_cell_size].uintval);
        return ret;
    }

    [[nodiscard]] hist_t histogramMax() const
    {
        uint64_t sum = 0, mx = 0;
        if (ffr_cuda_histogram_sum_max(ctx.get(),&sum,&mx) != FFR_OK)
            fail("BufferRenderer::histogramMax()");
        return mx;
    }

    [[nodiscard]] const ffr_flame_desc &getFlameDesc() const { return *desc; }
    [[nodiscard]] ffr_ctx *getContext() const { return ctx.get(); }
    [[nodiscard]] const std::vector<buf_elem_t> &getBuffer() const { return fetch(); }
    [[nodiscard]] const buf_elem_t *getBufferCell(size_t i) const { return fetch().data() + i*buffer_cell_size; }
    [[nodiscard]] size_t getBufferNumCells() const { return buffer_cells; }
    [[nodiscard]] size_t getBufferCellSize() const { return buffer_cell_size; }
    [[nodiscard]] size_t getSamplesIterated() const { return stats.s_iter; }
    [[nodiscard]] size_t getSamplesPlotted() const { return stats.s_plot; }
    [[nodiscard]] const std::vector<size_t> &getXFormDistribution() const { return stats.xf_dist; }
    [[nodiscard]] const std::vector<size_t> &getBadValueXForms() const { return stats.bv_xfs; }
    [[nodiscard]] const std::vector<point_t> &getBadValuePoints() const { return stats.bv_pts; }
    [[nodiscard]] const std::array<num_pair_t,dims> &getPointExtremes() const { return stats.pt_max; }
    [[nodiscard]] const std::array<num_t,dims> &getDimMults() const { return mult_d; }
    [[nodiscard]] const std::array<size_t,dims> &getIndexMults() const { return mult_i; }
    [[nodiscard]] size_t getDims() const { return dims; }
    [[nodiscard]] size_t getColorDims() const { return desc->color_dims; }
};

} // namespace ffr

#endif
