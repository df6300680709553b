/*
ffr_buf.cpp -- ffr-buf.out: the reference's buffer renderer command line over libffr_cuda.

Same flags, same buffer file format, same stderr report as the reference's src/ffr_buf.cpp
(flags :64-73, batch heuristic :94-101, report :104-119 and :226-256, -i handling :168-183,
output :259-272), so -i buffers still add up and ffr-img.out still reads the output. Boost is
replaced by getopt_long. Additive flags: --seed (the reference seeds from the clock, so a run
without --seed does the same), --gpus.

options:
-h,--help         show usage help message
-f,--flame        flame json file (required), - for stdin
-i,--input        input buffers to add at start (repeatable), - for stdin
-o,--output       output file (required), - for stdout
-s,--samples      number of samples to render
-t,--threads      accepted for compatibility; the GPU path does not use host threads
-b,--batch-size   samples per chain (work unit); 0 = calculate a reasonable size
-z,--bad-values   number of bad values allowed before terminating render
   --seed         base seed of the per-chain ISAAC streams (default: clock)
   --gpus         number of GPUs to shard the chains over (default 1): one process, the library's
                  multi-device context (private buffers, peer reduce-scatter at the end).
                  FFR_MULTI_PROCESS=1 runs one PROCESS per GPU instead (fork before any CUDA call;
                  the workers render their chain ranges into private buffers, the first process
                  adds them to its own over NVLink peer memory through CUDA IPC and writes the
                  output). Measured on the 8 x B200 box: the driver initialises ~0.7 s per visible
                  GPU and process, serialised system-wide, so eight processes (10.4-11.5 s to a
                  ready context) lose to one (7.1 s, of which 5.9 s are cuInit itself); on two GPUs
                  the two forms are level (2.2 s vs 2.7 s). Hence the default.
   --float        the reference's float/uint32_t build (types.hpp:24-41) instead of the shipped
                  double/uint64_t one: 4-byte buffer elements, ISAAC-32
*/

#include "../../include/ffr_flame.h"

#include <getopt.h>
#include <signal.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <fstream>
#include <iostream>
#include <iterator>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

static const std::string VERSION = "ffr-b200 0.1";

static void usage()
{
    std::cerr <<
        "ffr-buf usage:\n"
        "  -h [ --help ]                 show this help message\n"
        "  -f [ --flame ] arg            flame parameters JSON file\n"
        "  -i [ --input ] arg            buffers to add initially\n"
        "  -o [ --output ] arg           output file\n"
        "  -s [ --samples ] arg (=0)     samples to render\n"
        "  -t [ --threads ] arg          threads to use (ignored on the GPU path)\n"
        "  -b [ --batch-size ] arg (=0)  batch size (samples per chain)\n"
        "  -z [ --bad-values ] arg (=256) bad values limit\n"
        "  --seed arg                    base seed (default: from the clock)\n"
        "  --gpus arg (=1)               GPUs to use\n"
        "  --float                       float/uint32_t build (4-byte buffer elements)\n"
        "  --jit / --no-jit              compile a kernel for this flame at start-up (NVRTC) / never;\n"
        "                                default: only for renders large enough to repay the compile\n";
}

static bool read_all(std::istream& is, std::string& out)
{
    out.assign(std::istreambuf_iterator<char>(is),std::istreambuf_iterator<char>());
    return !is.bad();
}

struct Progress
{
    struct timespec t1;
    uint64_t samples, batch;
};

static void progress_cb(void *user, uint64_t done, uint64_t total)
{
    // same line as ffr_buf.cpp:196-213
    Progress *p = (Progress*)user;
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC,&t);
    size_t tn = 1000000000uLL * (t.tv_sec - p->t1.tv_sec) + (t.tv_nsec - p->t1.tv_nsec);
    size_t tavg = done ? tn / done : 0;
    size_t trem = tavg * (total - done);
    double frac = done / (double) total;
    std::cerr << "\rprogress: " << done << "/" << total << " batches, "
        << std::min<uint64_t>(p->samples,done*p->batch) << "/" << p->samples << " samples ("
        << (int)(100*frac) << "%) " << (tn / 1e9) << "sec elapsed, (~"
        << (trem / 1e9) << "sec remaining)";
}

/* ---- one process per GPU (--gpus N) ---- */

struct WorkerMsg
{
    int rc;                         /* FFR_OK / FFR_BAD_VALUES / < 0 */
    char err[256];
    unsigned char handle[FFR_IPC_HANDLE_BYTES];
    ffr_stats stats;
};

static bool write_all(int fd, const void *p, size_t n)
{
    const char *c = (const char*)p;
    while (n)
    {
        ssize_t w = write(fd,c,n);
        if (w <= 0)
            return false;
        c += w;
        n -= (size_t)w;
    }
    return true;
}

static bool read_all_fd(int fd, void *p, size_t n)
{
    char *c = (char*)p;
    while (n)
    {
        ssize_t r = read(fd,c,n);
        if (r <= 0)
            return false;
        c += r;
        n -= (size_t)r;
    }
    return true;
}

/* contiguous chain range of process `rank` (the rule ffr_cuda_render_chains uses between devices) */
static void chain_range(uint64_t chains, int nproc, int rank, uint64_t *first, uint64_t *count)
{
    const uint64_t per = (chains + nproc - 1) / nproc;
    *first = std::min<uint64_t>(per*rank,chains);
    *count = std::min<uint64_t>(per,chains - *first);
}

/* a worker: render the range, hand the buffer to the collector, wait until it is done with it */
[[noreturn]] static void worker_main(int rank, int nproc, const ffr_flame_desc *desc, const ffr_options &opt0,
        uint64_t chains, uint64_t chain_len, uint64_t last_len, uint64_t seed, uint64_t bv_limit,
        int up_fd, int down_fd)
{
    static WorkerMsg msg;
    memset(&msg,0,sizeof(msg));
    ffr_options opt = opt0;
    /* FFR_WORKER_MASK=1: the worker sees only its own GPU (CUDA_VISIBLE_DEVICES, entry `rank` of an
       inherited list), so that its driver initialisation does not touch the other devices */
    int dev = rank;
    if (getenv("FFR_WORKER_MASK") && *getenv("FFR_WORKER_MASK") == '1')
    {
        std::string mine = std::to_string(rank);
        if (const char *vis = getenv("CUDA_VISIBLE_DEVICES"))
        {
            std::stringstream ss(vis);
            std::string item;
            for (int k = 0; std::getline(ss,item,','); ++k)
                if (k == rank)
                    mine = item;
        }
        setenv("CUDA_VISIBLE_DEVICES",mine.c_str(),1);
        dev = 0;
    }
    ffr_ctx *ctx = ffr_cuda_create_ex(desc,&dev,1,&opt,msg.err,sizeof(msg.err));
    if (!ctx && opt.jit == 2)
    {
        opt.jit = 1;
        ctx = ffr_cuda_create_ex(desc,&dev,1,&opt,msg.err,sizeof(msg.err));
    }
    if (!ctx)
        msg.rc = FFR_E_CUDA;
    else
    {
        uint64_t first, count;
        chain_range(chains,nproc,rank,&first,&count);
        const bool has_last = first + count == chains;
        msg.rc = count ? ffr_cuda_render_chains(ctx,first,count,chain_len,has_last ? last_len : 0,seed,bv_limit,&msg.stats)
                       : ffr_cuda_get_stats(ctx,&msg.stats);
        if (msg.rc < 0)
            snprintf(msg.err,sizeof(msg.err),"%s",ffr_cuda_last_error(ctx));
        else if (ffr_cuda_ipc_export(ctx,msg.handle) != FFR_OK)
        {
            msg.rc = FFR_E_CUDA;
            snprintf(msg.err,sizeof(msg.err),"%s",ffr_cuda_last_error(ctx));
        }
    }
    write_all(up_fd,&msg,sizeof(msg));
    char done;
    read_all_fd(down_fd,&done,1);      /* the collector has added this buffer (or gave up) */
    if (ctx)
        ffr_cuda_destroy(ctx);
    _exit(0);
}

static void merge_stats(ffr_stats &a, const ffr_stats &b, uint32_t dims)
{
    a.s_iter += b.s_iter;
    a.s_plot += b.s_plot;
    for (int i = 0; i < FFR_MAX_XFORMS; ++i)
        a.xf_dist[i] += b.xf_dist[i];
    for (uint32_t d = 0; d < dims; ++d)
    {
        a.pt_min[d] = std::min(a.pt_min[d],b.pt_min[d]);
        a.pt_max[d] = std::max(a.pt_max[d],b.pt_max[d]);
    }
    const uint64_t na = std::min<uint64_t>(a.n_bad,FFR_MAX_BAD_RECORDED), nb = std::min<uint64_t>(b.n_bad,FFR_MAX_BAD_RECORDED);
    for (uint64_t k = 0; k < nb && na + k < FFR_MAX_BAD_RECORDED; ++k)
    {
        a.bad_xf[na + k] = b.bad_xf[k];
        for (int d = 0; d < FFR_MAX_DIMS; ++d)
            a.bad_pt[na + k][d] = b.bad_pt[k][d];
    }
    a.n_bad += b.n_bad;
}

int main(int argc, char **argv)
{
    size_t arg_threads = std::thread::hardware_concurrency();
    std::string arg_flame, arg_output;
    std::vector<std::string> arg_input;
    size_t arg_samples = 0, arg_batch_size = 0, arg_bad_values = 1 << 8;
    uint64_t arg_seed = 0;
    bool have_seed = false;
    int arg_gpus = 1;
    int arg_elem = 8;
    uint32_t arg_jit = 0;

    static const struct option longopts[] = {
        {"help",no_argument,nullptr,'h'},
        {"flame",required_argument,nullptr,'f'},
        {"input",required_argument,nullptr,'i'},
        {"output",required_argument,nullptr,'o'},
        {"samples",required_argument,nullptr,'s'},
        {"threads",required_argument,nullptr,'t'},
        {"batch-size",required_argument,nullptr,'b'},
        {"bad-values",required_argument,nullptr,'z'},
        {"seed",required_argument,nullptr,1000},
        {"gpus",required_argument,nullptr,1001},
        {"float",no_argument,nullptr,1002},
        {"jit",no_argument,nullptr,1003},
        {"no-jit",no_argument,nullptr,1004},
        {nullptr,0,nullptr,0}
    };
    if (argc < 2)
    {
        usage();
        return 1;
    }
    int c;
    while ((c = getopt_long(argc,argv,"hf:i:o:s:t:b:z:",longopts,nullptr)) != -1)
    {
        switch (c)
        {
        case 'h': usage(); return 1;
        case 'f': arg_flame = optarg; break;
        case 'i': arg_input.push_back(optarg); break;
        case 'o': arg_output = optarg; break;
        case 's': arg_samples = strtoull(optarg,nullptr,10); break;
        case 't': arg_threads = strtoull(optarg,nullptr,10); break;
        case 'b': arg_batch_size = strtoull(optarg,nullptr,10); break;
        case 'z': arg_bad_values = strtoull(optarg,nullptr,10); break;
        case 1000: arg_seed = strtoull(optarg,nullptr,0); have_seed = true; break;
        case 1001: arg_gpus = atoi(optarg); break;
        case 1002: arg_elem = 4; break;
        case 1003: arg_jit = 2; break;
        case 1004: arg_jit = 1; break;
        default: usage(); return 1;
        }
    }
    if (arg_flame.empty() || arg_output.empty())
    {
        std::cerr << "the options '--flame' and '--output' are required" << std::endl;
        return 1;
    }
    if (arg_gpus < 1)
        arg_gpus = 1;

    // batch size when 0: the reference's clamp((samples+255)>>8, 4096, 1<<20) (ffr_buf.cpp:94-101)
    // is sized for tens of host threads; a B200 runs ~75k chains at once, so the calculated
    // size aims at >= 8 chain groups per resident block, still >= the reference's minimum
    bool calculated_batch_size = false;
    if (arg_batch_size == 0)
    {
        uint64_t resident = (uint64_t)arg_gpus * 148 * 2 * 256;
        uint64_t guess = (arg_samples + resident*8 - 1) / (resident*8);
        arg_batch_size = std::clamp<uint64_t>(guess,1<<11,1<<16);
        calculated_batch_size = true;
    }
    if (!have_seed)
    {
        // Isaac::setSeed() (isaac.hpp:244-249) seeds from the clock; so does a run without --seed
        struct timespec t;
        clock_gettime(CLOCK_REALTIME,&t);
        arg_seed = (uint64_t)t.tv_sec * 1000000000uLL + t.tv_nsec;
    }

    std::cerr << "ffr-buf version " << VERSION << std::endl;
    std::cerr << "--flame " << arg_flame << std::endl;
    for (auto& s : arg_input)
        std::cerr << "--input " << s << std::endl;
    std::cerr << "--output " << arg_output << std::endl;
    std::cerr << "--samples " << arg_samples;
    if (arg_samples == 0)
        std::cerr << " (not rendering)";
    std::cerr << std::endl;
    std::cerr << "--threads " << arg_threads << " (unused: " << arg_gpus << " GPU)" << std::endl;
    if (calculated_batch_size)
        std::cerr << "--batch_size 0 (" << arg_batch_size << ")" << std::endl;
    else
        std::cerr << "--batch_size " << arg_batch_size << std::endl;
    std::cerr << "--bad_values " << arg_bad_values << std::endl;
    std::cerr << "--seed " << arg_seed << std::endl;
    std::cerr << "--" << std::endl;

    std::string text;
    if (arg_flame == "-")
    {
        if (!read_all(std::cin,text))
            throw std::runtime_error("error reading flame");
    }
    else
    {
        std::ifstream json_file(arg_flame,std::ios::in|std::ios::binary);
        if (!json_file || !read_all(json_file,text))
        {
            std::cerr << "ERROR: cannot read " << arg_flame << std::endl;
            return 1;
        }
    }
    /* FFR_TIMING=1: wall-clock of the phases on stderr (additive; off by default so that the
       report keeps the reference's lines) */
    const bool timing = getenv("FFR_TIMING") && *getenv("FFR_TIMING") == '1';
    const auto t_start = std::chrono::steady_clock::now();
    auto phase = [&](const char *name)
    {
        if (timing)
            std::cerr << "timing: " << name << " at "
                      << std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count()
                      << " s" << std::endl;
    };
    char err[512];
    {
        /* the reference echoes the parsed flame before anything else (ffr_buf.cpp:129) */
        err[0] = 0;
        const size_t n = ffr_flame_json_echo(text.data(),text.size(),nullptr,0,err,sizeof(err));
        if (!n)
        {
            std::cerr << "ERROR: " << err << std::endl;
            return 1;
        }
        std::string echo(n + 1,'\0');
        ffr_flame_json_echo(text.data(),text.size(),&echo[0],echo.size(),nullptr,0);
        echo.resize(n);
        std::cerr << "flame: " << echo << std::endl;
    }
    ffr_flame *flame = ffr_flame_from_json_ex(text.data(),text.size(),nullptr,0,arg_elem,err,sizeof(err));
    if (!flame)
    {
        std::cerr << "ERROR: " << err << std::endl;
        return 1;
    }
    const ffr_flame_desc *desc = ffr_flame_get_desc(flame);

    ffr_options opt;
    memset(&opt,0,sizeof(opt));
    opt.struct_size = sizeof(opt);
    opt.jit = arg_jit;

    /* the chains of the whole job (every process and device derives its share from these) */
    const uint64_t total_chains = arg_samples ? (arg_samples + arg_batch_size - 1) / arg_batch_size : 0;
    const uint64_t last_chain = arg_samples - (total_chains ? (total_chains - 1)*arg_batch_size : 0);
    const uint64_t last_len = (last_chain == arg_batch_size) ? 0 : last_chain;

    /* --gpus N: one process per GPU. Nothing up to this point has touched CUDA or NVRTC, so forking
       is safe. */
    const bool multi = getenv("FFR_MULTI_PROCESS") && *getenv("FFR_MULTI_PROCESS") == '1';
    const int nproc = (arg_gpus > 1 && multi && arg_samples > 0) ? arg_gpus : 1;
    std::vector<pid_t> pids;
    std::vector<int> up_fds, down_fds;
    if (nproc > 1)
    {
        if (arg_samples < 256*(uint64_t)1 || arg_batch_size < 256)
        {
            std::cerr << "ERROR: BufferRenderer::render(): batch size too small" << std::endl;
            return 1;
        }
        if (opt.jit == 0 && arg_samples >= 50000000000ull)
            opt.jit = 2;        /* the auto threshold applies to the job, not to one process's share */
        /* No NVRTC here: a process that has compiled cannot hand a usable CUDA state to its forked
           children (measured: the workers then find no device). Every process compiles for itself,
           all at once -- the wall-clock cost of one compile -- unless the kernel is in the disk
           cache already. */
        signal(SIGPIPE,SIG_IGN);
        for (int r = 1; r < nproc; ++r)
        {
            int up[2], down[2];
            if (pipe(up) != 0 || pipe(down) != 0)
            {
                std::cerr << "ERROR: pipe() failed" << std::endl;
                return 1;
            }
            const pid_t pid = fork();
            if (pid < 0)
            {
                std::cerr << "ERROR: fork() failed" << std::endl;
                return 1;
            }
            if (pid == 0)
            {
                close(up[0]);
                close(down[1]);
                for (int fd : up_fds) close(fd);
                for (int fd : down_fds) close(fd);
                worker_main(r,nproc,desc,opt,total_chains,arg_batch_size,last_len,arg_seed,arg_bad_values,
                    up[1],down[0]);
            }
            close(up[1]);
            close(down[0]);
            pids.push_back(pid);
            up_fds.push_back(up[0]);
            down_fds.push_back(down[1]);
        }
    }
    auto release_workers = [&]()
    {
        for (int fd : down_fds)
        {
            char done = 1;
            write_all(fd,&done,1);
            close(fd);
        }
        for (int fd : up_fds)
            close(fd);
        for (pid_t pid : pids)
            waitpid(pid,nullptr,0);
        down_fds.clear();
        up_fds.clear();
        pids.clear();
    };

    const int dev0 = 0;
    const int ctx_devices = nproc > 1 ? 1 : arg_gpus;
    ffr_ctx *ctx = ffr_cuda_create_ex(desc,nproc > 1 ? &dev0 : nullptr,ctx_devices,&opt,err,sizeof(err));
    if (!ctx && opt.jit == 2)
    {
        /* the flame-specialised kernel is an optimisation: fall back to the interpreter kernels */
        std::cerr << "note: " << err << "; using the ahead-of-time kernels" << std::endl;
        opt.jit = 1;
        ctx = ffr_cuda_create_ex(desc,nproc > 1 ? &dev0 : nullptr,ctx_devices,&opt,err,sizeof(err));
    }
    if (!ctx)
        release_workers();
    if (!ctx)
    {
        std::cerr << "ERROR: " << err << std::endl;
        return 1;
    }
    phase("context created");
    const size_t bytes = ffr_cuda_buffer_bytes(ctx);
    std::cerr << "buffer: " << bytes << " bytes, ";
#if __BYTE_ORDER__ == __ORDER_LITTLE_ENDIAN__
    std::cerr << "little endian, ";
#else
    std::cerr << "big endian, ";
#endif
    std::cerr << desc->elem_size << " byte numbers" << std::endl;
    std::cerr << "size:";
    for (uint32_t d = 0; d < desc->dims; ++d)
        std::cerr << " " << desc->size[d];
    std::cerr << std::endl;
    std::cerr << "color: " << desc->color_dims << " dimensions" << std::endl;

    std::vector<char> host(bytes);

    // add initial input buffers (ffr_buf.cpp:168-183)
    for (auto& s : arg_input)
    {
        std::cerr << "adding input file " << s << std::endl;
        bool ok;
        if (s == "-")
        {
            std::cin.read(host.data(),bytes);
            ok = std::cin.good() || (size_t)std::cin.gcount() == bytes;
        }
        else
        {
            std::ifstream input_file(s,std::ios::in|std::ios::binary);
            input_file.read(host.data(),bytes);
            ok = (size_t)input_file.gcount() == bytes;
        }
        if (!ok || ffr_cuda_add_buffer(ctx,host.data(),bytes) != FFR_OK)
        {
            std::cerr << "ERROR: error reading file" << std::endl;
            return 1;
        }
    }

    if (arg_samples > 0)
    {
        std::cerr << "render start" << std::endl;
        struct timespec t1,t2;
        Progress prog;
        clock_gettime(CLOCK_MONOTONIC,&t1);
        prog.t1 = t1;
        prog.samples = arg_samples;
        prog.batch = arg_batch_size;
        /* cb_thread of BufferRenderer::render (ffr_buf.cpp:214-218): one line per worker; the
           workers here are the devices */
        for (int g = 0; g < arg_gpus; ++g)
            std::cerr << "started thread " << g << " (cuda:" << g << ")" << std::endl;
        static ffr_stats stats;
        int rc;
        if (nproc == 1)
            rc = ffr_cuda_render(ctx,arg_samples,arg_batch_size,arg_seed,arg_bad_values,
                progress_cb,&prog,&stats);
        else
        {
            /* this process's share, in a few launches with a progress line after each (scaled to
               the job: the workers advance at the same rate), then the workers' buffers */
            uint64_t first, count;
            chain_range(total_chains,nproc,0,&first,&count);
            const uint64_t wave = std::max<uint64_t>(1,ffr_cuda_resident_chains(ctx));
            const uint64_t segs = std::min<uint64_t>(32,std::max<uint64_t>(1,count / (wave*4)));
            const uint64_t per = (count + segs - 1) / std::max<uint64_t>(1,segs);
            rc = FFR_OK;
            for (uint64_t f = 0; f < count && rc >= 0; f += per)
            {
                const uint64_t c = std::min<uint64_t>(per,count - f);
                const bool has_last = first + f + c == total_chains;
                const int r1 = ffr_cuda_render_chains(ctx,first + f,c,arg_batch_size,has_last ? last_len : 0,
                    arg_seed,arg_bad_values,&stats);
                rc = r1 < 0 ? r1 : std::max(rc,r1);
                progress_cb(&prog,std::min<uint64_t>(total_chains,(f + c)*nproc),total_chains);
                if (r1 == FFR_BAD_VALUES)
                    break;
            }
            static WorkerMsg msg;
            std::vector<unsigned char> handles;
            for (size_t w = 0; w < up_fds.size() && rc >= 0; ++w)
            {
                if (!read_all_fd(up_fds[w],&msg,sizeof(msg)))
                {
                    std::cerr << std::endl << "ERROR: worker " << (w + 1) << " died" << std::endl;
                    release_workers();
                    return 1;
                }
                if (msg.rc < 0)
                {
                    std::cerr << std::endl << "ERROR: worker " << (w + 1) << ": " << msg.err << std::endl;
                    release_workers();
                    return 1;
                }
                rc = std::max(rc,msg.rc);
                merge_stats(stats,msg.stats,desc->dims);
                handles.insert(handles.end(),msg.handle,msg.handle + FFR_IPC_HANDLE_BYTES);
            }
            if (rc >= 0 && !handles.empty() &&
                    ffr_cuda_ipc_add(ctx,handles.data(),(int)(handles.size()/FFR_IPC_HANDLE_BYTES)) != FFR_OK)
                rc = FFR_E_CUDA;
            release_workers();
            if (stats.n_bad > arg_bad_values && rc == FFR_OK)
                rc = FFR_BAD_VALUES;
        }
        // like the reference, the timer spans render() only (ffr_buf.cpp:189-227); the buffer
        // reduce + copy to the host happens below, as the reference's write does
        clock_gettime(CLOCK_MONOTONIC,&t2);
        std::cerr << std::endl;
        if (rc < 0)
        {
            std::cerr << "ERROR: " << ffr_cuda_last_error(ctx) << std::endl;
            return 1;
        }
        size_t nsecs = 1000000000uLL * (t2.tv_sec - t1.tv_sec) + (t2.tv_nsec - t1.tv_nsec);
        double secs = nsecs / 1e9;
        if (secs < 1e-9)
            secs = 1e-9;
        std::cerr << "render done: " << secs << " sec ("
            << (size_t)(arg_samples/secs) << " samples/sec)" << std::endl;
        if (rc == FFR_BAD_VALUES)
            std::cerr << "render failure "
                << " (flame may not be sufficiently contractive)" << std::endl;
        double iter_part = stats.s_iter / (double) arg_samples;
        std::cerr << "samples iterated: " << stats.s_iter
            << " (" << (100*iter_part) << "%)" << std::endl;
        double plot_part = stats.s_plot / (double) arg_samples;
        std::cerr << "samples plotted: " << stats.s_plot
            << " (" << (100*plot_part) << "%)" << std::endl;
        std::cerr << "xform selection:";
        for (uint32_t i = 0; i < desc->num_xform_ids; ++i)
            std::cerr << " " << stats.xf_dist[i];
        std::cerr << std::endl;
        uint64_t nb = std::min<uint64_t>(stats.n_bad,FFR_MAX_BAD_RECORDED);
        std::cerr << "bad value xforms:";
        for (uint64_t i = 0; i < nb; ++i)
            std::cerr << " " << stats.bad_xf[i];
        std::cerr << std::endl;
        std::cerr << "bad value points:";
        for (uint64_t i = 0; i < nb; ++i)
        {
            std::cerr << " (" << stats.bad_pt[i][0];
            for (uint32_t d = 1; d < desc->dims; ++d)
                std::cerr << "," << stats.bad_pt[i][d];
            std::cerr << ")";
        }
        std::cerr << std::endl;
        std::cerr << "extreme coordinates:";
        for (uint32_t d = 0; d < desc->dims; ++d)
            std::cerr << " (" << stats.pt_min[d] << "," << stats.pt_max[d] << ")";
        std::cerr << std::endl;
    }

    phase("render and report done");
    std::cerr << "writing output" << std::endl;
    if (ffr_cuda_read_buffer(ctx,host.data(),bytes) != FFR_OK)
    {
        std::cerr << "ERROR: " << ffr_cuda_last_error(ctx) << std::endl;
        return 1;
    }
    phase("buffer reduced and read back");
    if (arg_output == "-")
    {
        std::cout.write(host.data(),bytes);
        if (!std::cout.good())
        {
            std::cerr << "ERROR: error writing output" << std::endl;
            return 1;
        }
    }
    else
    {
        std::ofstream output_file(arg_output,std::ios::out|std::ios::binary);
        output_file.write(host.data(),bytes);
        if (!output_file.good())
        {
            std::cerr << "ERROR: error writing output" << std::endl;
            return 1;
        }
    }
    phase("output written");
    ffr_cuda_destroy(ctx);
    ffr_flame_free(flame);
    phase("context destroyed");
    /* everything is written and closed: leave without the CUDA runtime's exit handlers, which take
       ~0.35 s per device to unwind the primary contexts one after the other (8 devices: 2.8 s of an
       11.4 s run); the operating system reclaims the devices either way. FFR_CLEAN_EXIT=1 returns
       normally. */
    std::cerr.flush();
    std::cout.flush();
    if (!(getenv("FFR_CLEAN_EXIT") && *getenv("FFR_CLEAN_EXIT") == '1'))
        _exit(0);
    return 0;
}
