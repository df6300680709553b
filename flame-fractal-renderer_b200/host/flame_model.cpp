/*
flame_model.cpp -- host side flame model: flame JSON -> ffr_flame_desc (ffr_flame.h).

Restates, for the device path, what the reference does once at construction time.
Every derived parameter is computed on the host with the expression the reference
constructor uses (same operand order, same libm), so the device sees the identical
doubles. Citations are to the reference repo.
*/

#include "../../include/ffr_flame.h"
#include "json_min.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace
{

using ffr::Json;
using ffr::JsonError;

// per-precision constants, types/constants.hpp:17-37. The reference is compiled for one
// num_t (types.hpp:24-41); every constructor below runs in num_t arithmetic, so the whole
// model is templated on T = num_t and the resulting values (exactly representable in a double)
// are stored in the double fields of the POD desc.
template <typename T> struct Consts;
template <> struct Consts<double>
{
    static constexpr double eps = 1e-20;     // eps_v<double> :21
    static constexpr double max_rect = 1e10; // max_rect_v<double> :36
};
template <> struct Consts<float>
{
    static constexpr float eps = 1e-10;      // eps_v<float> :19
    static constexpr float max_rect = 1e5F;  // max_rect_v<float> :35
};
const uint64_t MAX_DIM = 65535;      // max_dim, constants.hpp:33

// math::sincosg, utils/math.hpp:21-24
inline void sincosg(float a, float& s, float& c) { sincosf(a,&s,&c); }
inline void sincosg(double a, double& s, double& c) { sincos(a,&s,&c); }

struct VarName { uint32_t op; const char *name; };

// factory names (variations.hpp:2387-2612)
const VarName VAR_NAMES[] = {
    {FFR_VAR_LINEAR,"linear"},{FFR_VAR_SINUSOIDAL,"sinusoidal"},
    {FFR_VAR_SPHERICAL,"spherical"},{FFR_VAR_BENT,"bent"},
    {FFR_VAR_RECTANGLES,"rectangles"},{FFR_VAR_FISHEYE,"fisheye"},
    {FFR_VAR_BUBBLE,"bubble"},{FFR_VAR_NOISE,"noise"},{FFR_VAR_BLUR,"blur"},
    {FFR_VAR_GAUSSIAN_BLUR,"gaussian_blur"},{FFR_VAR_SQUARE_NOISE,"square_noise"},
    {FFR_VAR_SEPARATION,"separation"},{FFR_VAR_SPLITS,"splits"},
    {FFR_VAR_PRE_BLUR,"pre_blur"},{FFR_VAR_MODULUS,"modulus"},{FFR_VAR_CELLN,"celln"},
    {FFR_VAR_SWIRL,"swirl"},{FFR_VAR_HORSESHOE,"horseshoe"},{FFR_VAR_POLAR,"polar"},
    {FFR_VAR_POLAR2,"polar2"},{FFR_VAR_HANDKERCHIEF,"handkerchief"},
    {FFR_VAR_HEART,"heart"},{FFR_VAR_DISC,"disc"},{FFR_VAR_DISC2,"disc2"},
    {FFR_VAR_WAVES,"waves"},{FFR_VAR_FAN,"fan"},{FFR_VAR_RINGS,"rings"},
    {FFR_VAR_SPIRAL,"spiral"},{FFR_VAR_HYPERBOLIC,"hyperbolic"},
    {FFR_VAR_DIAMOND,"diamond"},{FFR_VAR_EX,"ex"},{FFR_VAR_JULIA,"julia"},
    {FFR_VAR_EXPONENTIAL,"exponential"},{FFR_VAR_POWER,"power"},
    {FFR_VAR_COSINE,"cosine"},{FFR_VAR_BLOB,"blob"},{FFR_VAR_PDJ,"pdj"},
    {FFR_VAR_CYLINDER,"cylinder"},{FFR_VAR_PERSPECTIVE,"perspective"},
    {FFR_VAR_JULIAN,"julian"},{FFR_VAR_JULIASCOPE,"juliascope"},
    {FFR_VAR_RADIAL_BLUR,"radial_blur"},{FFR_VAR_PIE,"pie"},{FFR_VAR_NGON,"ngon"},
    {FFR_VAR_CURL,"curl"},{FFR_VAR_ARCH,"arch"},{FFR_VAR_TANGENT,"tangent"},
    {FFR_VAR_RAYS,"rays"},{FFR_VAR_BLADE,"blade"},{FFR_VAR_SECANT,"secant"},
    {FFR_VAR_TWINTRIAN,"twintrian"},{FFR_VAR_CROSS,"cross"},{FFR_VAR_EXP,"exp"},
    {FFR_VAR_LOG,"log"},{FFR_VAR_SIN,"sin"},{FFR_VAR_COS,"cos"},{FFR_VAR_TAN,"tan"},
    {FFR_VAR_SEC,"sec"},{FFR_VAR_CSC,"csc"},{FFR_VAR_COT,"cot"},{FFR_VAR_SINH,"sinh"},
    {FFR_VAR_COSH,"cosh"},{FFR_VAR_TANH,"tanh"},{FFR_VAR_SECH,"sech"},
    {FFR_VAR_CSCH,"csch"},{FFR_VAR_COTH,"coth"},{FFR_VAR_AUGER,"auger"},
    {FFR_VAR_FLUX,"flux"},{FFR_VAR_MOBIUS,"mobius"},{FFR_VAR_SCRY,"scry"},
    {FFR_VAR_SPLIT,"split"},{FFR_VAR_STRIPES,"stripes"},{FFR_VAR_WEDGE,"wedge"},
    {FFR_VAR_WEDGE_JULIA,"wedge_julia"},{FFR_VAR_WEDGE_SPH,"wedge_sph"},
    {FFR_VAR_WHORL,"whorl"},{FFR_VAR_SUPERSHAPE,"supershape"},{FFR_VAR_FLOWER,"flower"},
    {FFR_VAR_CONIC,"conic"},{FFR_VAR_PARABOLA,"parabola"},{FFR_VAR_BIPOLAR,"bipolar"},
    {FFR_VAR_BOARDERS,"boarders"},{FFR_VAR_BUTTERFLY,"butterfly"},{FFR_VAR_CELL,"cell"},
    {FFR_VAR_CPOW,"cpow"},{FFR_VAR_CURVE,"curve"},{FFR_VAR_EDISC,"edisc"},
    {FFR_VAR_ELLIPTIC,"elliptic"},{FFR_VAR_ESCHER,"escher"},{FFR_VAR_FOCI,"foci"},
    {FFR_VAR_LAZYSUSAN,"lazysusan"},{FFR_VAR_LOONIE,"loonie"},{FFR_VAR_OSCOPE,"oscope"},
    {FFR_VAR_POPCORN,"popcorn"},{FFR_VAR_SPHERICAL_P,"spherical_p"},
    {FFR_VAR_UNIT_SPHERE,"unit_sphere"},{FFR_VAR_UNIT_SPHERE_P,"unit_sphere_p"},
    {FFR_VAR_UNIT_CUBE,"unit_cube"},
};

bool is2d(uint32_t op) { return op >= FFR_VAR_FIRST_2D && op <= FFR_VAR_LAST_2D; }

// Point<T,N>(const Json&), types/point.hpp:59-77
template <typename T>
void parsePoint(const Json& j, uint32_t n, T *out)
{
    if (!j.isArray())
        throw JsonError("Point(Json&): not an array");
    const auto& a = j.arrayValue();
    if (a.size() != n)
        throw JsonError("Point(Json&): incorrect array size: "
            + std::to_string(a.size()));
    for (uint32_t i = 0; i < n; ++i)
    {
        if (a[i].isInt())
            out[i] = (T)a[i].intValue();
        else if (a[i].isFloat())
            out[i] = (T)a[i].floatValue();
        else
            throw JsonError("Point(Json&): entry is not a number");
    }
}

// Affine<T,N>::Affine(const Json&), types/affine.hpp:45-92
template <typename T>
void parseAffine(const Json& j, uint32_t n, double *Aout, double *bout)
{
    if (!j.isObject())
        throw JsonError("Affine(Json&): not an object");
    T A[9], b[3];
    for (uint32_t i = 0; i < 9; ++i) A[i] = 0;
    for (uint32_t i = 0; i < 3; ++i) b[i] = 0;
    struct Store
    {
        T *A, *b; double *Ao, *bo;
        ~Store() { for (int i = 0; i < 9; ++i) Ao[i] = A[i]; for (int i = 0; i < 3; ++i) bo[i] = b[i]; }
    } store{A,b,Aout,bout};
    if (!j.has("A"))
    {
        for (uint32_t i = 0; i < n; ++i)
            A[i*n+i] = 1;
    }
    else
    {
        const Json& ja = j["A"];
        if (!ja.isArray())
            throw JsonError("Affine(Json&): A is not an array");
        if (ja.size() != n)
            throw JsonError("Affine(Json&): A is wrong size");
        for (uint32_t i = 0; i < n; ++i)
        {
            try
            {
                parsePoint<T>(ja[i],n,A+i*n);
            }
            catch (const std::exception& e)
            {
                throw JsonError("Affine(Json&): error parsing A["
                    + std::to_string(i) + "]: " + e.what());
            }
        }
    }
    if (j.has("b"))
    {
        try
        {
            parsePoint<T>(j["b"],n,b);
        }
        catch (const std::exception& e)
        {
            throw JsonError("Affine(Json&): error parsing b: "
                + std::string(e.what()));
        }
    }
}

void identityAffine(uint32_t n, double *A, double *b)
{
    for (uint32_t i = 0; i < 9; ++i) A[i] = 0.0;
    for (uint32_t i = 0; i < 3; ++i) b[i] = 0.0;
    for (uint32_t i = 0; i < n; ++i)
        A[i*n+i] = 1;
}

inline double F(const Json& j, const char *key) { return j[key].floatValue(); }

// Variation<dims>::parseVariation and every constructor (variations.hpp)
template <typename T>
ffr_variation parseVariation(const Json& j, uint32_t D)
{
    const T EPS = Consts<T>::eps;
    std::string name = j["name"].stringValue();
    uint32_t op = ffr_var_op_from_name(name.c_str());
    // 2-d factory is empty below 2 dimensions (variations.hpp:2379-2384)
    if (op == 0 || (is2d(op) && D < 2))
        throw std::runtime_error("unknown variation: "+name);
    ffr_variation v;
    memset(&v,0,sizeof(v));
    v.op = op;
    v.weight = (T)j["weight"].floatValue(); // Variation ctor :45-48 (num_t weight)
    T p[FFR_MAX_VAR_PARAMS];
    for (int i = 0; i < FFR_MAX_VAR_PARAMS; ++i) p[i] = 0;
    // derived members are num_t; they are copied to the desc's double fields after the switch
    if (is2d(op) && D > 2) // VariationFrom2D ctor :72-88
    {
        int64_t ax = j["axis_x"].intValue();
        int64_t ay = j["axis_y"].intValue();
        if (ax < 0 || ax >= (int64_t)D)
            throw std::runtime_error("axis_x index out of range");
        if (ay < 0 || ay >= (int64_t)D)
            throw std::runtime_error("axis_y index out of range");
        if (ax == ay)
            throw std::runtime_error("axes are not distinct");
        v.axis_x = (uint32_t)ax;
        v.axis_y = (uint32_t)ay;
    }
    else
    {
        v.axis_x = 0;
        v.axis_y = 1;
    }
    switch (op)
    {
    case FFR_VAR_BENT: // :221-226
        parsePoint<T>(j["scales_neg"],D,p);
        parsePoint<T>(j["scales_pos"],D,p+4);
        break;
    case FFR_VAR_RECTANGLES: // :249-252
        parsePoint<T>(j["params"],D,p);
        break;
    case FFR_VAR_FISHEYE: // :279-284
    case FFR_VAR_BUBBLE:  // :301-306
        p[0] = F(j,"addval");
        break;
    case FFR_VAR_SEPARATION: // :387-393
        parsePoint<T>(j["params"],D,p);
        for (uint32_t i = 0; i < D; ++i)
            p[i] *= p[i];
        parsePoint<T>(j["inside"],D,p+4);
        break;
    case FFR_VAR_SPLITS: // :414-417
        parsePoint<T>(j["params"],D,p);
        break;
    case FFR_VAR_MODULUS: // :455-460
        parsePoint<T>(j["params"],D,p);
        for (uint32_t i = 0; i < D; ++i)
            p[i] *= 2.0;
        for (uint32_t i = 0; i < D; ++i)
            p[4+i] = 1.0/p[i];
        break;
    case FFR_VAR_CELLN: // :480-485
        parsePoint<T>(j["sizes"],D,p);
        for (uint32_t i = 0; i < D; ++i)
            p[4+i] = 1.0 / p[i];
        break;
    case FFR_VAR_DISC2: // :629-643
    {
        T rot = F(j,"rotation");
        T twist = F(j,"twist");
        p[0] = rot*M_PI;
        T sinadd,cosadd;
        sincosg((T)(twist),sinadd,cosadd);
        cosadd -= 1.0;
        T k = 1.0 + twist;
        if (twist > 2.0*M_PI) k -= 2.0*M_PI;
        if (twist < -2.0*M_PI) k += 2.0*M_PI;
        cosadd *= k;
        sinadd *= k;
        p[1] = cosadd;
        p[2] = sinadd;
        break;
    }
    case FFR_VAR_WAVES: // :664-670
        p[0] = F(j,"xfreq");
        p[1] = F(j,"xscale");
        p[2] = F(j,"yfreq");
        p[3] = F(j,"yscale");
        break;
    case FFR_VAR_FAN: // :689-695
    {
        T x = F(j,"x");
        T y = F(j,"y");
        p[0] = M_PI * (x*x + EPS);
        p[1] = y;
        break;
    }
    case FFR_VAR_RINGS: // :718-722
    {
        T val = F(j,"value");
        p[0] = val*val + EPS;
        break;
    }
    case FFR_VAR_BLOB: // :879-886
    {
        T low = F(j,"low");
        T high = F(j,"high");
        p[0] = (high+low)/2.0;
        p[1] = (high-low)/2.0;
        p[2] = F(j,"waves");
        break;
    }
    case FFR_VAR_PDJ: // :905-911
        p[0] = F(j,"a");
        p[1] = F(j,"b");
        p[2] = F(j,"c");
        p[3] = F(j,"d");
        break;
    case FFR_VAR_PERSPECTIVE: // :947-953
    {
        p[0] = F(j,"distance");
        T angle = F(j,"angle");
        p[1] = sin(angle);
        p[2] = p[0]*cos(angle);
        break;
    }
    case FFR_VAR_JULIAN:     // :971-978
    case FFR_VAR_JULIASCOPE: // :998-1005
    {
        T power = F(j,"power");
        T dist = F(j,"dist");
        p[0] = fabs(power);
        p[1] = 1.0/power;
        p[2] = dist/(2.0*power);
        break;
    }
    case FFR_VAR_RADIAL_BLUR: // :1027-1032
    {
        T angle = F(j,"angle");
        sincosg((T)(angle*M_PI_2),p[0],p[1]);
        p[2] = F(j,"flam3_weight");
        break;
    }
    case FFR_VAR_PIE: // :1054-1060
        p[0] = F(j,"slices");
        p[1] = F(j,"rotation");
        p[2] = F(j,"thickness");
        p[3] = (2.0*M_PI)/p[0];
        break;
    case FFR_VAR_NGON: // :1081-1089
    {
        T sides = F(j,"sides");
        p[0] = F(j,"power")/2.0;
        p[1] = (2.0*M_PI)/sides;
        p[2] = F(j,"corners");
        p[3] = F(j,"circle");
        p[4] = sides/(2.0*M_PI);
        break;
    }
    case FFR_VAR_CURL: // :1111-1115
        p[0] = F(j,"c1");
        p[1] = F(j,"c2");
        break;
    case FFR_VAR_ARCH:      // :1138-1141
    case FFR_VAR_RAYS:      // :1176-1179
    case FFR_VAR_BLADE:     // :1200-1203
    case FFR_VAR_SECANT:    // :1222-1225
    case FFR_VAR_TWINTRIAN: // :1244-1247
    case FFR_VAR_SCRY:      // :1639-1642
        p[0] = F(j,"flam3_weight");
        break;
    case FFR_VAR_AUGER: // :1553-1559
        p[0] = F(j,"freq");
        p[1] = F(j,"flam3_weight");
        p[2] = F(j,"scale") / 2.0;
        p[3] = F(j,"sym");
        break;
    case FFR_VAR_FLUX: // :1580-1585
        p[0] = 2.0 + F(j,"spread");
        p[1] = F(j,"flam3_weight");
        break;
    case FFR_VAR_MOBIUS: // :1609-1615
        parsePoint<T>(j["a"],2,p);
        parsePoint<T>(j["b"],2,p+2);
        parsePoint<T>(j["c"],2,p+4);
        parsePoint<T>(j["d"],2,p+6);
        break;
    case FFR_VAR_SPLIT: // :1659-1663
        p[0] = F(j,"xsize") * M_PI;
        p[1] = F(j,"ysize") * M_PI;
        break;
    case FFR_VAR_STRIPES: // :1682-1686
        p[0] = 1.0 - F(j,"space");
        p[1] = F(j,"warp");
        break;
    case FFR_VAR_WEDGE: // :1705-1712
        p[0] = F(j,"swirl");
        p[1] = F(j,"count");
        p[2] = F(j,"angle");
        p[3] = F(j,"hole");
        p[4] = 1.0 - p[2]*p[1]*(M_1_PI*0.5);
        break;
    case FFR_VAR_WEDGE_JULIA: // :1733-1743
    {
        T angle = F(j,"angle");
        T count = F(j,"count");
        T power = F(j,"power");
        T invpower = 1.0/power;
        T dist = F(j,"dist");
        p[0] = dist/(2.0*power);
        p[1] = fabs(power);
        p[2] = invpower;
        p[3] = count;
        p[4] = angle;
        p[5] = 1.0 - angle*count*(M_1_PI*0.5);
        break;
    }
    case FFR_VAR_WEDGE_SPH: // :1765-1772
    {
        T angle = F(j,"angle");
        T count = F(j,"count");
        p[0] = F(j,"swirl");
        p[1] = count;
        p[3] = angle;
        p[4] = F(j,"hole");
        p[2] = 1.0 - angle*count*(M_1_PI*0.5);
        break;
    }
    case FFR_VAR_WHORL: // :1794-1799
        p[0] = F(j,"inside");
        p[1] = F(j,"outside");
        p[2] = F(j,"flam3_weight");
        break;
    case FFR_VAR_SUPERSHAPE: // :1819-1828
    {
        T n1 = F(j,"n1");
        p[0] = F(j,"m") / 4.0;
        p[1] = -1.0 / n1;
        p[2] = F(j,"n2");
        p[3] = F(j,"n3");
        p[4] = F(j,"rnd");
        p[5] = F(j,"holes");
        break;
    }
    case FFR_VAR_FLOWER: // :1851-1855
        p[0] = F(j,"petals");
        p[1] = F(j,"holes");
        break;
    case FFR_VAR_CONIC: // :1873-1877
        p[0] = F(j,"eccen");
        p[1] = F(j,"holes");
        break;
    case FFR_VAR_PARABOLA: // :1895-1899
        p[0] = F(j,"height");
        p[1] = F(j,"width");
        break;
    case FFR_VAR_BIPOLAR: // :1918-1921
        p[0] = -M_PI_2 * F(j,"shift");
        break;
    case FFR_VAR_BOARDERS: // :1944-1950
        p[0] = F(j,"prob");
        if (p[0] < 0.0 || p[0] > 1.0)
            throw std::runtime_error("boarders probability out of range");
        break;
    case FFR_VAR_CELL: // :2012-2016
        p[0] = F(j,"size");
        p[1] = 1.0/p[0];
        break;
    case FFR_VAR_CPOW: // :2042-2050
    {
        T r = F(j,"r");
        T i = F(j,"i");
        T power = F(j,"power");
        p[3] = power;
        p[0] = 2.0*M_PI/power;
        p[1] = r/power;
        p[2] = i/power;
        break;
    }
    case FFR_VAR_CURVE: // :2070-2081
    {
        p[2] = F(j,"xamp");
        p[3] = F(j,"yamp");
        T xlen = F(j,"xlen");
        T ylen = F(j,"ylen");
        xlen *= xlen;
        ylen *= ylen;
        p[0] = 1.0 / std::max(EPS,xlen);
        p[1] = 1.0 / std::max(EPS,ylen);
        break;
    }
    case FFR_VAR_ESCHER: // :2153-2160
    {
        T beta = F(j,"beta");
        T seb,ceb;
        sincosg((T)(beta),seb,ceb);
        p[0] = 0.5*(1.0+ceb);
        p[1] = 0.5*seb;
        break;
    }
    case FFR_VAR_LAZYSUSAN: // :2201-2209
        p[0] = F(j,"x");
        p[1] = F(j,"y");
        p[2] = F(j,"spin");
        p[3] = F(j,"twist");
        p[4] = F(j,"space");
        p[5] = F(j,"flam3_weight");
        break;
    case FFR_VAR_LOONIE: // :2239-2243
        p[0] = F(j,"flam3_weight");
        p[1] = p[0] * p[0];
        break;
    case FFR_VAR_OSCOPE: // :2263-2270
    {
        T freq = F(j,"frequency");
        p[0] = 2.0*M_PI*freq;
        p[1] = F(j,"amplitude");
        p[2] = F(j,"damping");
        p[3] = F(j,"separation");
        break;
    }
    case FFR_VAR_POPCORN: // :2290-2295
        p[0] = F(j,"x");
        p[1] = F(j,"y");
        p[2] = F(j,"c");
        break;
    case FFR_VAR_SPHERICAL_P:   // :2316-2321
    case FFR_VAR_UNIT_SPHERE_P: // :2351-2356
        p[0] = F(j,"norm");
        if (p[0] <= 0.0)
            throw std::runtime_error("norm <= 0");
        break;
    default: // parameterless
        break;
    }
    for (int i = 0; i < FFR_MAX_VAR_PARAMS; ++i)
        v.params[i] = p[i];
    return v;
}

struct XFormModel
{
    ffr_xform x;
    std::vector<ffr_variation> vars;
    std::vector<double> color;
};

// XForm<dims>::XForm, types/xform.hpp:71-172
template <typename T>
XFormModel parseXForm(const Json& in, uint64_t id, bool is_final, uint32_t D,
        uint32_t color_dims, T default_color_speed)
{
    XFormModel m;
    memset(&m.x,0,sizeof(m.x));
    m.x.id = id;
    if (!is_final)
    {
        try
        {
            m.x.weight = (T)in["weight"].floatValue();
        }
        catch (std::exception& e)
        {
            throw JsonError("XForm(): cannot parse weight: " + std::string(e.what()));
        }
    }
    else
        m.x.weight = 1.0; // unused
    if (m.x.weight < 0.0)
        throw JsonError("XForm(): weight is negative");
    Json affine;
    m.x.has_pre = in.valueAt("pre_affine",affine);
    if (m.x.has_pre)
    {
        try
        {
            parseAffine<T>(affine,D,m.x.pre_A,m.x.pre_b);
        }
        catch (std::exception& e)
        {
            throw JsonError("XForm(): cannot parse pre_affine" + std::string(e.what()));
        }
    }
    else
        identityAffine(D,m.x.pre_A,m.x.pre_b);
    m.x.has_post = in.valueAt("post_affine",affine);
    if (m.x.has_post)
    {
        try
        {
            parseAffine<T>(affine,D,m.x.post_A,m.x.post_b);
        }
        catch (std::exception& e)
        {
            throw JsonError("XForm(): cannot parse post_affine" + std::string(e.what()));
        }
    }
    else
        identityAffine(D,m.x.post_A,m.x.post_b);
    try
    {
        for (const Json& varj : in["variations"].arrayValue())
            m.vars.push_back(parseVariation<T>(varj,D));
    }
    catch (std::exception& e)
    {
        throw JsonError("XForm(): cannot parse variations: " + std::string(e.what()));
    }
    Json colorj;
    if (color_dims && in.valueAt("color",colorj))
    {
        try
        {
            const auto& ca = colorj.arrayValue();
            if (ca.size() != color_dims)
                throw JsonError("color length incorrect");
            m.color.resize(color_dims);
            for (uint32_t i = 0; i < color_dims; ++i)
            {
                m.color[i] = (T)ca[i].floatValue();
                if (m.color[i] < 0.0 || m.color[i] > 1.0)
                    throw JsonError("color coordinate out of range");
            }
        }
        catch (std::exception& e)
        {
            throw JsonError("XForm(): cannot parse color: " + std::string(e.what()));
        }
    }
    if (color_dims && in.valueAt("color_speed",colorj))
    {
        try
        {
            m.x.color_speed = (T)colorj.floatValue();
            if (m.x.color_speed < 0.0 || m.x.color_speed > 1.0)
                throw JsonError("color speed out of range");
        }
        catch (std::exception& e)
        {
            throw JsonError("XForm(): color speed issue: " + std::string(e.what()));
        }
    }
    else
        m.x.color_speed = default_color_speed;
    // XForm::_optimize, xform.hpp:47-58: drop zero weight variations
    m.vars.erase(std::remove_if(m.vars.begin(),m.vars.end(),
        [](const ffr_variation& v){ return fabs(v.weight) == 0.0; }),m.vars.end());
    m.x.has_color = !m.color.empty();
    m.x.num_vars = (uint32_t)m.vars.size();
    return m;
}

} // namespace

struct ffr_flame
{
    ffr_flame_desc desc;
    std::vector<XFormModel> models;   // selection order
    XFormModel final_model;
    std::vector<ffr_xform> xforms;    // desc.xforms
    std::vector<double> xfcw;

    void link()
    {
        xforms.clear();
        for (auto& m : models)
        {
            m.x.vars = m.vars.empty() ? nullptr : m.vars.data();
            m.x.color = m.color.empty() ? nullptr : m.color.data();
            xforms.push_back(m.x);
        }
        final_model.x.vars = final_model.vars.empty() ? nullptr : final_model.vars.data();
        final_model.x.color = final_model.color.empty() ? nullptr : final_model.color.data();
        desc.num_xforms = (uint32_t)xforms.size();
        desc.xforms = xforms.data();
        desc.xfcw = xfcw.data();
        desc.final_xform = desc.has_final ? &final_model.x : nullptr;
    }
};

namespace
{

// Flame<dims>::Flame(const Json&), types/flame.hpp:91-210
template <typename T>
ffr_flame *buildFlame(const Json& input)
{
    std::unique_ptr<ffr_flame> fl(new ffr_flame);
    ffr_flame_desc& d = fl->desc;
    memset(&d,0,sizeof(d));
    // ffr_buf.cpp:131-142
    int64_t dims = input["dimensions"].intValue();
    if (dims < 1 || dims > 3)
        throw JsonError(std::to_string(dims) + "D not supported");
    uint32_t D = (uint32_t)dims;
    d.dims = D;
    d.elem_size = (uint32_t)sizeof(T);
    try
    {
        const auto& sizej = input["size"].arrayValue();
        const auto& boundsj = input["bounds"].arrayValue();
        if (sizej.size() != D)
            throw JsonError("incorrect size length");
        if (boundsj.size() != D)
            throw JsonError("incorrect bounds length");
        T M = Consts<T>::max_rect;
        for (uint32_t i = 0; i < D; ++i)
        {
            try
            {
                // size[i] = sizej[i].floatValue(): double narrowed to size_t, flame.hpp:106
                d.size[i] = (uint64_t)sizej[i].floatValue();
            }
            catch (std::exception& e)
            {
                throw JsonError("error parsing size[" + std::to_string(i) + "]: " + e.what());
            }
            if (d.size[i] == 0 || d.size[i] > MAX_DIM)
                throw JsonError("size[" + std::to_string(i) + "] out of range");
            const auto& boundj = boundsj[i].arrayValue();
            if (boundj.size() != 2)
                throw JsonError("bounds[" + std::to_string(i) + "] wrong format");
            T lo,hi;
            try
            {
                lo = (T)boundj[0].floatValue();
                hi = (T)boundj[1].floatValue();
            }
            catch (std::exception& e)
            {
                throw JsonError("error parsing bounds[" + std::to_string(i) + "]: " + e.what());
            }
            d.bounds_lo[i] = lo;
            d.bounds_hi[i] = hi;
            if (lo < -M || lo > M || hi < -M || hi > M)
                throw JsonError("bounds[" + std::to_string(i) + "] out of range");
            if (lo >= hi)
                throw JsonError("bounds[" + std::to_string(i) + "] low >= high");
        }
    }
    catch (std::exception& e)
    {
        throw JsonError("Flame(): " + std::string(e.what()));
    }
    Json fxf;
    d.has_final = input.valueAt("final_xform",fxf);
    const std::vector<Json> *xfs;
    try
    {
        xfs = &input["xforms"].arrayValue();
    }
    catch (std::exception& e)
    {
        throw JsonError("Flame(): cannot parse xforms: " + std::string(e.what()));
    }
    uint64_t color_dims;
    T color_speed;
    Json cd,cs;
    try
    {
        if (input.valueAt("color_dimensions",cd))
            color_dims = (uint64_t)cd.intValue();
        else
            color_dims = 0;
        if (color_dims > 127)
            throw JsonError("too many color dimensions");
        if (input.valueAt("color_speed",cs))
            color_speed = (T)cs.floatValue();
        else
            color_speed = 0.5;
        if (color_speed < 0.0 || color_speed > 1.0)
            throw std::runtime_error("color speed out of range");
    }
    catch (std::exception& e)
    {
        throw JsonError("Flame(): " + std::string(e.what()));
    }
    d.color_dims = (uint32_t)color_dims;
    uint64_t id = 0;
    for (const Json& xf : *xfs)
    {
        try
        {
            fl->models.push_back(parseXForm<T>(xf,id,false,D,d.color_dims,color_speed));
        }
        catch (std::exception& e)
        {
            throw JsonError("Flame(): error parsing xforms[" + std::to_string(id) + "]: " + e.what());
        }
        ++id;
    }
    d.num_xform_ids = (uint32_t)id;
    if (fl->models.empty())
        throw JsonError("Flame(): no xforms");
    if (d.has_final)
    {
        try
        {
            fl->final_model = parseXForm<T>(fxf,FFR_FINAL_XFORM_ID,true,D,d.color_dims,color_speed);
        }
        catch (std::exception& e)
        {
            throw JsonError("Flame(): error parsing final xform: " + std::string(e.what()));
        }
    }
    // Flame::_optimize, flame.hpp:63-83
    fl->models.erase(std::remove_if(fl->models.begin(),fl->models.end(),
        [](const XFormModel& m){ return m.x.weight == 0.0; }),fl->models.end());
    if (fl->models.empty())
        throw JsonError("Flame(): no xforms remaining after optimization");
    // same std::sort + comparator as flame.hpp:80-82 so the permutation (which depends only
    // on the sequence of comparison outcomes) is the one the reference gets
    std::sort(fl->models.begin(),fl->models.end(),
        [](XFormModel& a, XFormModel& b){ return a.x.weight > b.x.weight; });
    // Flame::_setupCumulativeWeights, flame.hpp:44-60
    size_t k = fl->models.size();
    fl->xfcw.assign(k,0.0);
    T normdiv = 0.0;
    for (size_t i = 0; i < k; ++i)
        normdiv += (T)fl->models[i].x.weight;
    T wsum = 0.0;
    for (size_t i = 0; i < k; ++i)
    {
        wsum += (T)fl->models[i].x.weight / normdiv;
        fl->xfcw[i] = wsum;
    }
    fl->xfcw.back() = 1.0;
    fl->link();
    return fl.release();
}

void setErr(char *err, size_t errlen, const std::string& msg)
{
    if (err && errlen)
    {
        snprintf(err,errlen,"%s",msg.c_str());
    }
}

} // namespace

extern "C"
{

ffr_flame *ffr_flame_from_json_ex(const char *text, size_t len, const uint64_t *size,
        int n_size, int elem_size, char *err, size_t errlen)
{
    try
    {
        if (elem_size != 8 && elem_size != 4)
            throw JsonError("elem_size must be 8 (double/uint64_t) or 4 (float/uint32_t)");
        Json j = Json::parse(text,len);
        if (size)
        {
            int64_t dims = j["dimensions"].intValue();
            if (dims != n_size)
                throw JsonError("size override length does not match dimensions");
            j.set("size",Json::makeArrayOfInts(size,(size_t)n_size));
        }
        return elem_size == 8 ? buildFlame<double>(j) : buildFlame<float>(j);
    }
    catch (std::exception& e)
    {
        setErr(err,errlen,e.what());
        return nullptr;
    }
}

ffr_flame *ffr_flame_from_json_sized(const char *text, size_t len, const uint64_t *size,
        int n_size, char *err, size_t errlen)
{
    return ffr_flame_from_json_ex(text,len,size,n_size,8,err,errlen);
}

ffr_flame *ffr_flame_from_json(const char *text, size_t len, char *err, size_t errlen)
{
    return ffr_flame_from_json_ex(text,len,nullptr,0,8,err,errlen);
}

const ffr_flame_desc *ffr_flame_get_desc(const ffr_flame *f)
{
    return f ? &f->desc : nullptr;
}

void ffr_flame_free(ffr_flame *f)
{
    delete f;
}

int ffr_flame_layout(const ffr_flame_desc *desc, double mult_d[FFR_MAX_DIMS],
        uint64_t mult_i[FFR_MAX_DIMS], uint64_t *cells, uint64_t *cell_size)
{
    // BufferRenderer::_init, buffer_renderer.hpp:114-140, in num_t arithmetic;
    // scale_adjust_down_v<num_t> = 1 - emach (constants.hpp:25-30,59-62)
    const bool f32 = desc->elem_size == 4;
    uint64_t n = 1;
    for (uint32_t i = 0; i < desc->dims; ++i)
    {
        uint64_t size = desc->size[i];
        double m;
        if (f32)
        {
            float mf = (float)(size) / ((float)desc->bounds_hi[i] - (float)desc->bounds_lo[i]);
            mf *= 1.0F - 1.0F / (float)(1 << 23);
            m = mf;
        }
        else
        {
            m = (double)(size) / (desc->bounds_hi[i] - desc->bounds_lo[i]);
            m *= 1.0 - (double)(float)(1.0 / (double)(1L << 52));
        }
        if (mult_d) mult_d[i] = m;
        if (mult_i) mult_i[i] = n;
        n *= size;
        if (n >= (1uLL << 48))
            return FFR_E_INVALID;
    }
    if (cells) *cells = n;
    if (cell_size) *cell_size = 1 + desc->color_dims;
    return FFR_OK;
}

const char *ffr_var_name(uint32_t op)
{
    for (const VarName& v : VAR_NAMES)
        if (v.op == op)
            return v.name;
    return nullptr;
}

uint32_t ffr_var_op_from_name(const char *name)
{
    for (const VarName& v : VAR_NAMES)
        if (!strcmp(v.name,name))
            return v.op;
    return 0;
}

size_t ffr_flame_json_echo(const char *text, size_t len, char *out, size_t outlen,
        char *err, size_t errlen)
{
    try
    {
        std::string d;
        Json::parse(text,len).dump(d);
        if (out && outlen)
            snprintf(out,outlen,"%s",d.c_str());
        return d.size();
    }
    catch (std::exception& e)
    {
        setErr(err,errlen,e.what());
        return 0;
    }
}

uint64_t ffr_reference_batch_size(uint64_t samples)
{
    uint64_t guess = (samples+255) >> 8;
    return std::clamp<uint64_t>(guess,1<<12,1<<20);
}

} // extern "C"
