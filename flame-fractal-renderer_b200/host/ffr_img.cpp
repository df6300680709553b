/*
ffr_img.cpp -- ffr-img.out: buffer(s) -> PNG with the tone map running on the GPU.

Same flags and the same pixel math as the reference's src/ffr_img.cpp (flags :67-77, checks
:86-95 and :123-127,:151-158, report :98-112,:217-230): the input buffers are summed on the
device (ffr_cuda_add_buffer), ffr_cuda_tonemap maps cells to pixels, and a small zlib based
PNG writer replaces boost::gil + libpng (absent here; the reference's image I/O is outside the
hot path). 2-d flames only, as in the reference.

-h,--help  -f,--flame  -i,--input (repeat)  -o,--output (png)  -y,--gamma
-m,--monochrome  -g,--grayscale  -c,--color (3 colour dims)  -b,--bits (8|16)
*/

#include "../../include/ffr_flame.h"

#include <getopt.h>
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <iterator>
#include <string>
#include <vector>

static const std::string VERSION = "ffr-b200 0.1";

static void put32(std::vector<unsigned char>& v, uint32_t x)
{
    v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x);
}

static void chunk(std::ostream& os, const char *type, const std::vector<unsigned char>& data)
{
    std::vector<unsigned char> hd;
    put32(hd,(uint32_t)data.size());
    os.write((const char*)hd.data(),4);
    std::vector<unsigned char> body(type,type+4);
    body.insert(body.end(),data.begin(),data.end());
    os.write((const char*)body.data(),body.size());
    std::vector<unsigned char> crc;
    put32(crc,(uint32_t)crc32(0L,body.data(),(uInt)body.size()));
    os.write((const char*)crc.data(),4);
}

/* pixels: h rows of w pixels, `channels` samples of `bits` bits in host byte order */
static bool write_png(std::ostream& os, const unsigned char *pixels, uint32_t w, uint32_t h,
        uint32_t channels, uint32_t bits)
{
    static const unsigned char sig[8] = {0x89,'P','N','G',0x0d,0x0a,0x1a,0x0a};
    os.write((const char*)sig,8);
    std::vector<unsigned char> ihdr;
    put32(ihdr,w);
    put32(ihdr,h);
    ihdr.push_back((unsigned char)bits);
    ihdr.push_back(channels == 3 ? 2 : 0); /* truecolour / grayscale */
    ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    chunk(os,"IHDR",ihdr);
    const size_t bps = bits/8, row = (size_t)w*channels*bps;
    std::vector<unsigned char> raw((row+1)*h);
    for (uint32_t y = 0; y < h; ++y)
    {
        unsigned char *dst = raw.data() + (row+1)*y;
        *dst++ = 0; /* filter: none */
        const unsigned char *src = pixels + row*y;
        if (bps == 1)
            memcpy(dst,src,row);
        else
            for (size_t i = 0; i < (size_t)w*channels; ++i) /* PNG samples are big endian */
            {
                uint16_t v;
                memcpy(&v,src+2*i,2);
                dst[2*i] = v >> 8;
                dst[2*i+1] = v & 0xff;
            }
    }
    uLongf clen = compressBound((uLong)raw.size());
    std::vector<unsigned char> comp(clen);
    if (compress2(comp.data(),&clen,raw.data(),(uLong)raw.size(),6) != Z_OK)
        return false;
    comp.resize(clen);
    chunk(os,"IDAT",comp);
    chunk(os,"IEND",{});
    return os.good();
}

static void usage()
{
    std::cerr <<
        "ffr-img usage:\n"
        "  -h [ --help ]            show this help message\n"
        "  -f [ --flame ] arg       flame parameters JSON file\n"
        "  -i [ --input ] arg       buffers to add\n"
        "  -o [ --output ] arg      output file\n"
        "  -y [ --gamma ] arg (=1)  gamma value\n"
        "  -m [ --monochrome ]      binary image from histogram\n"
        "  -g [ --grayscale ]       grayscale image from histogram\n"
        "  -c [ --color ]           color image from 3d color only\n"
        "  -b [ --bits ] arg (=8)   bits per color channel\n"
        "  --float                  input buffers are from the float/uint32_t build\n";
}

int main(int argc, char **argv)
{
    std::string arg_flame, arg_output;
    std::vector<std::string> arg_input;
    double arg_gamma = 1.0;
    size_t arg_bits = 8;
    bool arg_m = false, arg_g = false, arg_c = false;
    int arg_elem = 8; /* --float: buffers of the float/uint32_t build */
    static const struct option longopts[] = {
        {"help",no_argument,nullptr,'h'}, {"flame",required_argument,nullptr,'f'},
        {"input",required_argument,nullptr,'i'}, {"output",required_argument,nullptr,'o'},
        {"gamma",required_argument,nullptr,'y'}, {"monochrome",no_argument,nullptr,'m'},
        {"grayscale",no_argument,nullptr,'g'}, {"color",no_argument,nullptr,'c'},
        {"bits",required_argument,nullptr,'b'}, {"float",no_argument,nullptr,1002},
        {nullptr,0,nullptr,0}
    };
    if (argc < 2)
    {
        usage();
        return 1;
    }
    int c;
    while ((c = getopt_long(argc,argv,"hf:i:o:y:mgcb:",longopts,nullptr)) != -1)
    {
        switch (c)
        {
        case 'f': arg_flame = optarg; break;
        case 'i': arg_input.push_back(optarg); break;
        case 'o': arg_output = optarg; break;
        case 'y': arg_gamma = strtod(optarg,nullptr); break;
        case 'm': arg_m = true; break;
        case 'g': arg_g = true; break;
        case 'c': arg_c = true; break;
        case 'b': arg_bits = strtoull(optarg,nullptr,10); break;
        case 1002: arg_elem = 4; break;
        default: usage(); return 1;
        }
    }
    if (arg_flame.empty() || arg_output.empty())
    {
        std::cerr << "the options '--flame' and '--output' are required" << std::endl;
        return 1;
    }
    if (arg_gamma < 1e-20)
    {
        std::cerr << "ERROR: gamma too small" << std::endl;
        return 1;
    }
    if (arg_bits != 8 && arg_bits != 16)
    {
        std::cerr << "ERROR: bits per channel must be 8 or 16" << std::endl;
        return 1;
    }
    std::cerr << "ffr-img version " << VERSION << std::endl;
    std::cerr << "--flame " << arg_flame << std::endl;
    for (auto& s : arg_input)
        std::cerr << "--input " << s << std::endl;
    std::cerr << "--output " << arg_output << std::endl;
    std::cerr << "--gamma " << arg_gamma << std::endl;
    if (arg_m) std::cerr << "--monochrome" << std::endl;
    if (arg_g) std::cerr << "--grayscale" << std::endl;
    if (arg_c) std::cerr << "--color" << std::endl;
    std::cerr << "--bits " << arg_bits << std::endl;
    std::cerr << "--" << std::endl;

    std::string text;
    if (arg_flame == "-")
        text.assign(std::istreambuf_iterator<char>(std::cin),std::istreambuf_iterator<char>());
    else
    {
        std::ifstream f(arg_flame,std::ios::in|std::ios::binary);
        if (!f)
        {
            std::cerr << "ERROR: cannot read " << arg_flame << std::endl;
            return 1;
        }
        text.assign(std::istreambuf_iterator<char>(f),std::istreambuf_iterator<char>());
    }
    char err[512];
    ffr_flame *flame = ffr_flame_from_json_ex(text.data(),text.size(),nullptr,0,arg_elem,err,sizeof(err));
    if (!flame)
    {
        std::cerr << "ERROR: " << err << std::endl;
        return 1;
    }
    const ffr_flame_desc *desc = ffr_flame_get_desc(flame);
    if (desc->dims != 2)
    {
        std::cerr << "ERROR: only 2D flames supported" << std::endl;
        return 1;
    }
    ffr_ctx *ctx = ffr_cuda_create(desc,nullptr,1,err,sizeof(err));
    if (!ctx)
    {
        std::cerr << "ERROR: " << err << std::endl;
        return 1;
    }
    const size_t bytes = ffr_cuda_buffer_bytes(ctx);
    std::cerr << "buffer: " << bytes << " bytes, little endian, " << desc->elem_size
        << " byte numbers" << std::endl;
    std::cerr << "color: " << desc->color_dims << " dimensions" << std::endl;
    if ((int)arg_m + (int)arg_g + (int)arg_c != 1)
    {
        std::cerr << "ERROR: must choose exactly 1 coloring flag "
            << "(--monochrome, --grayscale, --color)" << std::endl;
        return 1;
    }
    if (arg_input.empty())
    {
        std::cerr << "ERROR: no input buffer files specified" << std::endl;
        return 1;
    }
    std::vector<char> host(bytes);
    for (auto& s : arg_input)
    {
        std::cerr << "adding input file " << s << std::endl;
        size_t got;
        if (s == "-")
        {
            std::cin.read(host.data(),bytes);
            got = (size_t)std::cin.gcount();
        }
        else
        {
            std::ifstream f(s,std::ios::in|std::ios::binary);
            f.read(host.data(),bytes);
            got = (size_t)f.gcount();
        }
        if (got != bytes || ffr_cuda_add_buffer(ctx,host.data(),bytes) != FFR_OK)
        {
            std::cerr << "ERROR: error reading file" << std::endl;
            return 1;
        }
    }
    std::cerr << "processing buffer" << std::endl;
    const int mode = arg_m ? FFR_TONE_MONO : (arg_g ? FFR_TONE_GRAY : FFR_TONE_RGB);
    const uint32_t channels = arg_c ? 3 : 1;
    const uint32_t bits = arg_m ? 8 : (uint32_t)arg_bits;
    std::vector<unsigned char> pixels((size_t)desc->size[0]*desc->size[1]*channels*(bits/8));
    ffr_tonemap_info info;
    int rc = ffr_cuda_tonemap(ctx,mode,(int)bits,arg_gamma,pixels.data(),pixels.size(),&info);
    if (rc != FFR_OK && std::string(ffr_cuda_last_error(ctx)) != "histogram is (probably) empty")
    {
        std::cerr << "ERROR: " << ffr_cuda_last_error(ctx) << std::endl;
        return 1;
    }
    std::cerr << "histogram bounds: " << info.hist_min << " " << info.hist_max << std::endl;
    std::cerr << "scaler bounds: " << info.scaler_min << " " << info.scaler_max << std::endl;
    if (rc != FFR_OK)
    {
        std::cerr << "ERROR: histogram is (probably) empty" << std::endl;
        return 1;
    }
    bool ok;
    if (arg_output == "-")
        ok = write_png(std::cout,pixels.data(),info.width,info.height,channels,bits);
    else
    {
        std::ofstream f(arg_output,std::ios::out|std::ios::binary);
        ok = write_png(f,pixels.data(),info.width,info.height,channels,bits);
    }
    if (!ok)
    {
        std::cerr << "ERROR: error writing output" << std::endl;
        return 1;
    }
    ffr_cuda_destroy(ctx);
    ffr_flame_free(flame);
    return 0;
}
