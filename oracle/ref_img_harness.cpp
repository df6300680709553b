/*
oracle/ref_img_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Pins the tone map and the "flame:" echo to the reference's own code.

The reference's ffr-img cannot be built here as a program (Boost.GIL, Boost.Program_options
and libpng are absent), but its pixel arithmetic can:
  * renderers/image_renderer.hpp (getValueBounds, renderGrayImage, renderColorImageRGB,
    :112-192) is #included UNMODIFIED from /root/reference/src, against the small stand-in for
    the GIL types it touches (oracle/stub_gil/boost/gil.hpp);
  * render_image() of ffr_img.cpp (:199-309: the histogram/log bounds, the gamma, the three
    colouring lambdas) is extracted VERBATIM at build time by oracle/Makefile
    (`sed -n '/^bool render_image(/,$p'`) into oracle/_ref/gen/render_image.inc -- a build
    product under the git-ignored oracle/_ref/, never committed -- and #included below after
    the few globals it reads (same names and types as ffr_img.cpp:46-52);
  * writePng (declared in utils/image.hpp:28-29, defined in the un-buildable utils/image.cpp)
    is defined here to hand the finished image's pixels back instead of encoding a PNG.
So every number that decides a pixel comes out of reference-compiled code.

  refimg_render      flame text + raw buffer -> pixels (mode 1 mono / 2 gray / 3 rgb, 8/16 bit)
  refimg_flame_echo  `os << Json` (utils/json.cpp:203-207), the text after "flame: "
*/

#include "renderers/image_renderer.hpp"

#include "utils/image.hpp"
#include "utils/json.hpp"

#include <cstdint>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

// the globals render_image() reads (ffr_img.cpp:46-52)
tkoz::flame::num_t arg_gamma = 1.0;
size_t arg_bits = 8;
bool arg_m,arg_g,arg_c;

namespace
{
thread_local std::string g_err;
std::vector<unsigned char> g_pixels;
size_t g_w = 0, g_h = 0, g_ch = 0, g_bytes_per_channel = 0;
}

namespace tkoz::flame
{
template <typename img_t>
bool writePng(const img_t& img, std::ostream& os)
{
    (void)os;
    g_w = img.width();
    g_h = img.height();
    g_ch = (size_t)img_t::channels;
    g_bytes_per_channel = sizeof(typename img_t::channel_t);
    g_pixels.resize(img.data.size()*sizeof(typename img_t::pixel_t));
    memcpy(g_pixels.data(),img.data.data(),g_pixels.size());
    return true;
}
}

bool render_image(std::ostream&,tkoz::flame::ImageRenderer<>&);
#include "render_image.inc"

extern "C"
{

const char *refimg_last_error() { return g_err.c_str(); }

int refimg_elem_size() { return (int)sizeof(tkoz::flame::hist_t); }

/* returns 0 and fills pixels (w*h*channels*bits/8 bytes, row-major like the PNG rows);
   info = {hist_min, hist_max} as printed by "histogram bounds:"; log_bounds = "scaler bounds:" */
int refimg_render(const char *flame_text, const void *buf, size_t bytes, int mode, int bits,
        double gamma, void *pixels, size_t pix_bytes, uint64_t *hist_bounds, double *log_bounds)
{
    using namespace tkoz::flame;
    std::streambuf *old = std::cerr.rdbuf();
    std::ostringstream captured;
    try
    {
        Json j{std::string(flame_text)};
        Flame<2> flame(j);
        ImageRenderer<> img_ren(flame);
        std::string raw((const char*)buf,bytes);
        std::istringstream is(raw);
        if (!img_ren.getBufferRenderer().addBuffer(is))
            throw std::runtime_error("error reading file");
        arg_gamma = (num_t)gamma;
        if (arg_gamma < eps_v<num_t>)               // ffr_img.cpp:88-89
            throw std::runtime_error("gamma too small");
        arg_bits = (size_t)bits;
        arg_m = mode == 1;
        arg_g = mode == 2;
        arg_c = mode == 3;
        std::ostringstream sink;
        std::cerr.rdbuf(captured.rdbuf());
        bool ok = render_image(sink,img_ren);
        std::cerr.rdbuf(old);
        if (!ok)
            throw std::runtime_error("render_image failed");
        if (g_pixels.size() != pix_bytes)
            throw std::runtime_error("pixel buffer size mismatch: " + std::to_string(g_pixels.size()));
        memcpy(pixels,g_pixels.data(),pix_bytes);
        // "histogram bounds: a b\nscaler bounds: c d\n"
        std::istringstream rep(captured.str());
        std::string w1,w2;
        unsigned long long a = 0, b = 0;
        double c = 0, d = 0;
        rep >> w1 >> w2 >> a >> b >> w1 >> w2 >> c >> d;
        if (hist_bounds) { hist_bounds[0] = a; hist_bounds[1] = b; }
        if (log_bounds) { log_bounds[0] = c; log_bounds[1] = d; }
        return 0;
    }
    catch (std::exception& e)
    {
        std::cerr.rdbuf(old);
        g_err = e.what();
        return -1;
    }
}

size_t refimg_flame_echo(const char *text, char *out, size_t outlen)
{
    try
    {
        tkoz::flame::Json j{std::string(text)};
        std::ostringstream os;
        os << j;
        const std::string s = os.str();
        if (out && outlen)
            snprintf(out,outlen,"%s",s.c_str());
        return s.size();
    }
    catch (std::exception& e)
    {
        g_err = e.what();
        return 0;
    }
}

} // extern "C"
