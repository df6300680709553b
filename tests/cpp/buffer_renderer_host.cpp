/*
buffer_renderer_host.cpp -- a C++ host written against ffr::BufferRenderer<dims>
(include/ffr_buffer_renderer.hpp) the way the reference's run_renderer<dims>() is written against
tkoz::flame::BufferRenderer<dims> (src/ffr_buf.cpp:146-274): construct from the flame JSON, add
input buffers, render with callbacks, write the buffer, read the statistics through the getters.
Built and driven by tests/test_cpp_mirror.py; prints one JSON object on stdout.

usage: buffer_renderer_host FLAME.json OUT.buf SAMPLES BATCH SEED SPLIT [IN.buf ...]
   SPLIT > 0: the samples are rendered by two calls, SPLIT first and the rest second
usage: buffer_renderer_host --construct FLAME.json      (constructor only; prints what it threw)
*/
#include <ffr_buffer_renderer.hpp>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>

static std::string slurp(const char *path)
{
    std::ifstream f(path,std::ios::binary);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

static std::string quote(const std::string &s)
{
    std::string o = "\"";
    for (char c : s)
    {
        if (c == '"' || c == '\\')
            o += '\\';
        o += (c == '\n') ? ' ' : c;
    }
    return o + "\"";
}

/* a double as JSON (Python's json reads Infinity / -Infinity: the extremes before the first sample) */
static std::string number(double v)
{
    if (std::isinf(v))
        return v > 0 ? "Infinity" : "-Infinity";
    char num[40];
    snprintf(num,sizeof num,"%.17g",v);
    return num;
}

template <typename F> static std::string thrown(F f)
{
    try { f(); }
    catch (const std::runtime_error &e) { return e.what(); }
    return "";
}

template <size_t dims>
static int run(const std::string &json, int argc, char **argv)
{
    typedef ffr::BufferRenderer<dims> renderer_t;
    const char *out_path = argv[2];
    const size_t samples = strtoull(argv[3],nullptr,10), batch = strtoull(argv[4],nullptr,10);
    const uint64_t seed = strtoull(argv[5],nullptr,10);
    const size_t split = strtoull(argv[6],nullptr,10);
    renderer_t renderer(json);
    bool added = true;
    for (int i = 7; i < argc; ++i)
    {
        std::ifstream in(argv[i],std::ios::binary);
        added = added && renderer.addBuffer(in);
    }
    /* argument checks of render(), buffer_renderer.hpp:281-289 */
    const std::string e_threads0 = thrown([&]{ renderer.render(1000,0,batch,10); });
    const std::string e_threads = thrown([&]{ renderer.render(1000,65536,batch,10); });
    const std::string e_batch = thrown([&]{ renderer.render(1000,4,255,10); });
    const std::string e_vec = thrown([&]{ renderer.addBuffer(std::vector<uint64_t>(3)); });
    size_t batches = 0, workers = 0;
    renderer.setSeed(seed);
    bool ok = true;
    if (split)
    {
        ok = renderer.render(split,4,batch,10,[&]{ ++batches; },[&](const std::thread&,size_t){ ++workers; });
        ok = ok && renderer.renderSeeded(samples - split,batch,10,[&]{ ++batches; });
    }
    else
        ok = renderer.render(samples,4,batch,10,[&]{ ++batches; },[&](const std::thread&,size_t){ ++workers; });
    {
        std::ofstream out(out_path,std::ios::binary);
        if (!renderer.writeBuffer(out))
            return 3;
    }
    /* a second renderer that takes the first one's buffer twice: pointer and renderer overloads */
    renderer_t twice(json);
    twice.addBuffer(renderer);
    twice.addBuffer(renderer.getBuffer().data());
    size_t cell_sum = 0;
    for (size_t i = 0; i < renderer.getBufferNumCells(); ++i)
        cell_sum += renderer.getBufferCell(i)->uintval;
    std::cout << "{\"ok\": " << (ok ? "true" : "false") << ", \"added\": " << (added ? "true" : "false")
        << ", \"batches\": " << batches << ", \"workers\": " << workers
        << ", \"s_iter\": " << renderer.getSamplesIterated() << ", \"s_plot\": " << renderer.getSamplesPlotted()
        << ", \"sum\": " << renderer.histogramSum() << ", \"min\": " << renderer.histogramMin()
        << ", \"max\": " << renderer.histogramMax() << ", \"cell_sum\": " << cell_sum
        << ", \"twice_sum\": " << twice.histogramSum()
        << ", \"cells\": " << renderer.getBufferNumCells() << ", \"cell_size\": " << renderer.getBufferCellSize()
        << ", \"dims\": " << renderer.getDims() << ", \"color_dims\": " << renderer.getColorDims()
        << ", \"bad\": " << renderer.getBadValueXForms().size() << ", \"bad_pts\": " << renderer.getBadValuePoints().size()
        << ", \"next_seed\": " << renderer.getSeed() << ", \"xf_dist\": [";
    for (size_t i = 0; i < renderer.getXFormDistribution().size(); ++i)
        std::cout << (i ? ", " : "") << renderer.getXFormDistribution()[i];
    std::cout << "], \"extremes\": [";
    for (size_t d = 0; d < dims; ++d)
        std::cout << (d ? ", " : "") << "[" << number(renderer.getPointExtremes()[d].first) << ", "
            << number(renderer.getPointExtremes()[d].second) << "]";
    std::cout << "], \"mult_d\": [";
    for (size_t d = 0; d < dims; ++d)
        std::cout << (d ? ", " : "") << number(renderer.getDimMults()[d]);
    std::cout << "], \"mult_i\": [";
    for (size_t d = 0; d < dims; ++d)
        std::cout << (d ? ", " : "") << renderer.getIndexMults()[d];
    std::cout << "], \"e_threads0\": " << quote(e_threads0) << ", \"e_threads\": " << quote(e_threads)
        << ", \"e_batch\": " << quote(e_batch) << ", \"e_vec\": " << quote(e_vec) << "}" << std::endl;
    return 0;
}

int main(int argc, char **argv)
{
    if (argc == 3 && std::string(argv[1]) == "--construct")
    {
        const std::string json = slurp(argv[2]);
        try
        {
            ffr::BufferRenderer<2> renderer(json);
            std::cout << "{\"constructed\": true}" << std::endl;
        }
        catch (const ffr::JsonError &e)
        {
            std::cout << "{\"constructed\": false, \"kind\": \"JsonError\", \"what\": " << quote(e.what()) << "}" << std::endl;
        }
        catch (const std::runtime_error &e)
        {
            std::cout << "{\"constructed\": false, \"kind\": \"runtime_error\", \"what\": " << quote(e.what()) << "}" << std::endl;
        }
        return 0;
    }
    if (argc < 7)
    {
        std::cerr << "usage: " << argv[0] << " FLAME.json OUT.buf SAMPLES BATCH SEED SPLIT [IN.buf ...]" << std::endl;
        return 2;
    }
    const std::string json = slurp(argv[1]);
    /* ffr_buf.cpp:134-142: the dimension count picks the instantiation */
    size_t dims = 2;
    {
        const size_t k = json.find("\"dimensions\"");
        if (k != std::string::npos)
            dims = strtoull(json.c_str() + json.find(':',k) + 1,nullptr,10);
    }
    try
    {
        switch (dims)
        {
        case 1: return run<1>(json,argc,argv);
        case 2: return run<2>(json,argc,argv);
        case 3: return run<3>(json,argc,argv);
        }
    }
    catch (const std::exception &e)
    {
        std::cerr << "error: " << e.what() << std::endl;
        return 1;
    }
    std::cerr << "unsupported dimensions" << std::endl;
    return 2;
}
