/*
oracle/ffr_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Plain C restatement of the reference's chaos-game iterate-and-accumulate path
(tkoz0/flame-fractal-renderer, double/u64 build). It is the checker for libffr_cuda:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may
load it. The product never links or calls it and has no CPU fallback.

Parity of THIS file is pinned, bit for bit, against the unmodified reference compiled
by oracle/Makefile into oracle/_ref/libffr_ref.so (tests/test_oracle_vs_reference.py,
run where /root/reference exists) and against the golden vectors generated from that
build (tests/golden/, tests/golden/make_golden.py). The reference itself ships no
tests, golden vectors or fixtures for this path.

Compile: gcc -std=c11 -O3 -DNDEBUG (no -march, no -ffast-math): x86-64 baseline has no
FMA, so every expression below is evaluated in strict IEEE double exactly as the
reference's Release build evaluates it. Expression ORDER is part of the contract; do
not "simplify" anything here. Citations are file:line in the reference repo (src/).

Input is the POD flatten (include/ffr_cuda.h) produced by the host flame model; the
derived variation parameters in it are the members the reference constructors store.
*/

#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "ffr_cuda.h"

typedef uint64_t u64;
typedef int32_t i32;

/* ---- constants: types/constants.hpp:17-72 ---- */
#define EPS 1e-20                      /* eps_v<double> :21 */
#define SETTLE_ITERS 53                /* settle_iters_v<double> :46 */
#define BAD_VALUE_THRESHOLD 1e20       /* bad_value_threshold_v<double> :55 */

/* utils/flame.hpp:26-29 */
static inline int bad_value(double n)
{
    return fabs(n) > BAD_VALUE_THRESHOLD || isnan(n);
}

/* ---- ISAAC-64, RANDSIZL = 4: rng/isaac.hpp ---- */
#define RANDSIZ 16
typedef struct
{
    u64 randcnt;
    u64 randrsl[RANDSIZ];
    u64 randmem[RANDSIZ];
    u64 randa, randb, randc;
} isaac64;

/* ind, isaac.hpp:140-143 */
#define IND(mm,x) ((mm)[((x) >> 3) & (RANDSIZ-1)])

/* rngstep, isaac.hpp:146-153 (rparam = 4) */
#define RNGSTEP(mix) do { \
    x = *m; \
    a = (mix) + *(m2++); \
    *(m++) = y = IND(mm,x) + a + b; \
    *(r++) = b = IND(mm,y >> 4) + x; \
} while (0)

/* gen, isaac.hpp:77-90 with rngstep4 (u64) :196-203 */
static void isaac_gen(isaac64 *s)
{
    u64 a,b,x,y,*m,*mm,*m2,*r,*mend;
    mm = s->randmem;
    r = s->randrsl;
    a = s->randa;
    b = s->randb + (++s->randc);
    for (m = mm, mend = m2 = m+(RANDSIZ/2); m < mend;)
    {
        RNGSTEP(~(a^(a<<21)));
        RNGSTEP(  a^(a>> 5) );
        RNGSTEP(  a^(a<<12) );
        RNGSTEP(  a^(a>>33) );
    }
    for (m2 = mm; m2 < mend;)
    {
        RNGSTEP(~(a^(a<<21)));
        RNGSTEP(  a^(a>> 5) );
        RNGSTEP(  a^(a<<12) );
        RNGSTEP(  a^(a>>33) );
    }
    s->randa = a;
    s->randb = b;
}

/* mix (u64), isaac.hpp:171-182 */
#define MIX(a,b,c,d,e,f,g,h) do { \
    a -= e; f ^= h >>  9; h += a; \
    b -= f; g ^= a <<  9; a += b; \
    c -= g; h ^= b >> 23; b += c; \
    d -= h; a ^= c << 15; c += d; \
    e -= a; b ^= d >> 14; d += e; \
    f -= b; c ^= e << 20; e += f; \
    g -= c; d ^= f >> 17; f += g; \
    h -= d; e ^= g << 14; g += h; \
} while (0)

/* setSeed(u64) :267-271 -> setSeed(a0,b0,c0) :274-282 -> init(flag=false) :93-131 */
static void isaac_seed(isaac64 *s, u64 seed)
{
    u64 a,b,c,d,e,f,g,h;
    u64 *m = s->randmem;
    int i;
    memset(s->randrsl,0,sizeof(s->randrsl));
    s->randa = seed;
    s->randb = ~seed;
    s->randc = seed ^ 11713835213681433683uLL;
    a = b = c = d = e = f = g = h = 0x9e3779b97f4a7c13uLL; /* golden_ratio<u64> :36 */
    MIX(a,b,c,d,e,f,g,h);
    MIX(a,b,c,d,e,f,g,h);
    MIX(a,b,c,d,e,f,g,h);
    MIX(a,b,c,d,e,f,g,h);
    for (i = 0; i < RANDSIZ; i += 8)
    {
        MIX(a,b,c,d,e,f,g,h);
        m[i+0] = a; m[i+1] = b; m[i+2] = c; m[i+3] = d;
        m[i+4] = e; m[i+5] = f; m[i+6] = g; m[i+7] = h;
    }
    isaac_gen(s);
    s->randcnt = RANDSIZ;
}

/* next, isaac.hpp:321-329: results consumed from index 15 down to 0 */
static inline u64 isaac_next(isaac64 *s)
{
    if (s->randcnt-- == 0)
    {
        isaac_gen(s);
        s->randcnt = RANDSIZ-1;
    }
    return s->randrsl[s->randcnt];
}

/* ---- FlameRNG<double,u64,4>: rng/flame_rng.hpp ---- */

/* randNum :84-85 */
static inline double rand_num(isaac64 *s)
{
    return (isaac_next(s) >> 11) / (double)(1LL << 53);
}

/* randBool :61-64 */
static inline int rand_bool(isaac64 *s)
{
    return isaac_next(s) & 1;
}

/* randGaussian :143-148 via randGaussianPair :115-126 (Box-Muller, z2 wasted) */
static inline double rand_gaussian(isaac64 *s)
{
    double u1 = rand_num(s);
    double u2 = (2.0*M_PI)*rand_num(s);
    double r = sqrt(-2.0*log(u1));
    double sn,cs;
    sincos(u2,&sn,&cs);
    return r*cs;
}

/* randDirection<1|2|3> :171-206 */
static inline void rand_direction(isaac64 *s, int D, double *dir)
{
    if (D == 1)
        dir[0] = copysign(1.0,rand_num(s)-0.5);
    else if (D == 2)
    {
        double a = (2.0*M_PI) * rand_num(s);
        double sa,ca;
        sincos(a,&sa,&ca);
        dir[0] = ca;
        dir[1] = sa;
    }
    else
    {
        double u = 2.0*rand_num(s) - 1.0;
        double t = (2.0*M_PI) * rand_num(s);
        double r = sqrt(1.0 - u*u);
        double st,ct;
        sincos(t,&st,&ct);
        dir[0] = r*ct;
        dir[1] = r*st;
        dir[2] = u;
    }
}

/* ---- Point norms: types/point.hpp ---- */

/* normsum<2> :316-320 */
static inline double norm2sq(const double *v, int D)
{
    double ret = v[0]*v[0];
    for (int i = 1; i < D; ++i)
        ret += v[i]*v[i];
    return ret;
}

/* norm<2> :271-288 (N == 1 returns |x|) */
static inline double norm2(const double *v, int D)
{
    if (D == 1)
        return fabs(v[0]);
    return sqrt(norm2sq(v,D));
}

/* norm<0> :271-281 */
static inline double norminf(const double *v, int D)
{
    double ret = fabs(v[0]);
    for (int i = 1; i < D; ++i)
    {
        double a = fabs(v[i]);
        ret = (ret < a) ? a : ret; /* std::max */
    }
    return ret;
}

/* normsum(T p) :327-333 */
static inline double normsum_p(const double *v, int D, double p)
{
    double ret = pow(fabs(v[0]),p);
    for (int i = 1; i < D; ++i)
        ret += pow(fabs(v[i]),p);
    return ret;
}

/* ---- variations: variations/variations.hpp ---- */

/* 2-d variation bodies (calc2d). x,y in; ox,oy out. */
static void calc2d(const ffr_variation *v, isaac64 *rng, double x, double y,
        double *ox, double *oy)
{
    const double *p = v->params;
    const double xy[2] = {x,y};
    switch (v->op)
    {
    case FFR_VAR_SWIRL: /* :513-521 */
    {
        double r = norm2sq(xy,2);
        double sr,cr;
        sincos(r,&sr,&cr);
        *ox = x*sr-y*cr;
        *oy = x*cr+y*sr;
        return;
    }
    case FFR_VAR_HORSESHOE: /* :531-539 */
    {
        double r = 1.0 / (norm2(xy,2) + EPS);
        *ox = ((x-y)*(x+y))*r;
        *oy = (2.0*x*y)*r;
        return;
    }
    case FFR_VAR_POLAR: /* :549-554 */
    {
        double a = atan2(y,x);
        double r = norm2(xy,2);
        *ox = a*M_1_PI;
        *oy = r-1.0;
        return;
    }
    case FFR_VAR_POLAR2: /* :564-568 */
        *ox = atan2(y,x);
        *oy = log(norm2sq(xy,2));
        return;
    case FFR_VAR_HANDKERCHIEF: /* :578-583 */
    {
        double a = atan2(y,x);
        double r = norm2(xy,2);
        *ox = sin(a+r)*r;
        *oy = cos(a-r)*r;
        return;
    }
    case FFR_VAR_HEART: /* :593-600 */
    {
        double a = atan2(y,x);
        double r = norm2(xy,2);
        double sa,ca;
        sincos(r*a,&sa,&ca);
        *ox = sa*r;
        *oy = (-ca)*r;
        return;
    }
    case FFR_VAR_DISC: /* :610-617 */
    {
        double a = atan2(y,x);
        double r = norm2(xy,2);
        double sr,cr;
        sincos(M_PI*r,&sr,&cr);
        *ox = sr*a;
        *oy = cr*a;
        return;
    }
    case FFR_VAR_DISC2: /* :644-653 */
    {
        double t = p[0] * (x + y);
        double st,ct;
        sincos(t,&st,&ct);
        double a = atan2(y,x);
        *ox = (ct + p[1])*a;
        *oy = (st + p[2])*a;
        return;
    }
    case FFR_VAR_WAVES: /* :671-678 */
    {
        double dx = p[1]*sin(y*p[0]);
        double dy = p[3]*sin(x*p[2]);
        *ox = x + dx;
        *oy = y + dy;
        return;
    }
    case FFR_VAR_FAN: /* :696-707 */
    {
        double dx = p[0], dy = p[1];
        double dx2 = dx*0.5;
        double a = atan2(y,x);
        double m = copysign(1.0,dx2-fmod(a+dy,dx));
        a += m*dx2;
        double sa,ca;
        sincos(a,&sa,&ca);
        double r = norm2(xy,2);
        *ox = ca*r;
        *oy = sa*r;
        return;
    }
    case FFR_VAR_RINGS: /* :723-731 */
    {
        double dx = p[0];
        double r = norm2(xy,2);
        double s = y / r;
        double c = x / r;
        r = fmod(r+dx,2.0*dx) - dx + r*(1.0-dx);
        *ox = c*r;
        *oy = s*r;
        return;
    }
    case FFR_VAR_SPIRAL: /* :741-750 */
    {
        double r = norm2(xy,2);
        double sa = y / r;
        double ca = x / r;
        double sr,cr;
        sincos(r,&sr,&cr);
        double r1 = 1.0 / (r + EPS);
        *ox = (ca+sr)*r1;
        *oy = (sa-cr)*r1;
        return;
    }
    case FFR_VAR_HYPERBOLIC: /* :760-765 */
    {
        double r = norm2(xy,2);
        double sa = y / r;
        double ca = x / r;
        *ox = sa/(r+EPS);
        *oy = ca*r;
        return;
    }
    case FFR_VAR_DIAMOND: /* :775-782 */
    {
        double r = norm2(xy,2);
        double sa = y / r;
        double ca = x / r;
        double sr,cr;
        sincos(r,&sr,&cr);
        *ox = sa*cr;
        *oy = ca*sr;
        return;
    }
    case FFR_VAR_EX: /* :792-801 */
    {
        double a = atan2(y,x);
        double r = norm2(xy,2);
        double n0 = sin(a+r);
        double n1 = cos(a-r);
        double m0 = n0*n0*n0 * r;
        double m1 = n1*n1*n1 * r;
        *ox = m0+m1;
        *oy = m0-m1;
        return;
    }
    case FFR_VAR_JULIA: /* :811-818 */
    {
        double a = 0.5*atan2(y,x) + rand_bool(rng)*M_PI;
        double sa,ca;
        sincos(a,&sa,&ca);
        double r = norm2(xy,2);
        *ox = ca*r;
        *oy = sa*r;
        return;
    }
    case FFR_VAR_EXPONENTIAL: /* :828-836 */
    {
        double dx = exp(x-1.0);
        double sdy,cdy;
        sincos(M_PI*y,&sdy,&cdy);
        *ox = cdy*dx;
        *oy = sdy*dx;
        return;
    }
    case FFR_VAR_POWER: /* :846-851 */
    {
        double r = norm2(xy,2);
        double sa = y / r;
        double ca = x / r;
        double pw = pow(r,sa);
        *ox = ca*pw;
        *oy = sa*pw;
        return;
    }
    case FFR_VAR_COSINE: /* :861-868 */
    {
        double sa,ca;
        sincos(x*M_PI,&sa,&ca);
        *ox = ca*cosh(y);
        *oy = -sa*sinh(y);
        return;
    }
    case FFR_VAR_BLOB: /* :887-894 */
    {
        double r = norm2(xy,2);
        double sa = y / r;
        double ca = x / r;
        double a = atan2(y,x);
        r *= p[0] + p[1]*sin(p[2]*a);
        *ox = ca*r;
        *oy = sa*r;
        return;
    }
    case FFR_VAR_PDJ: /* :912-921 */
    {
        double nx1 = cos(p[1]*x);
        double nx2 = sin(p[2]*x);
        double ny1 = sin(p[0]*y);
        double ny2 = cos(p[3]*y);
        *ox = ny1-nx1;
        *oy = nx2-ny2;
        return;
    }
    case FFR_VAR_CYLINDER: /* :933-936 */
        *ox = sin(x);
        *oy = y;
        return;
    case FFR_VAR_PERSPECTIVE: /* :954-960 */
    {
        double t = 1.0 / (p[0] - y*p[1]);
        *ox = (p[0]*x)*t;
        *oy = (p[2]*y)*t;
        return;
    }
    case FFR_VAR_JULIAN: /* :979-987 */
    {
        i32 t = trunc(p[0]*rand_num(rng));
        double a = (atan2(y,x) + (2.0*M_PI)*t) * p[1];
        double r = pow(norm2sq(xy,2),p[2]);
        double sa,ca;
        sincos(a,&sa,&ca);
        *ox = ca*r;
        *oy = sa*r;
        return;
    }
    case FFR_VAR_JULIASCOPE: /* :1006-1015 */
    {
        i32 t = trunc(p[0]*rand_num(rng));
        double dir = copysign(1.0,rand_num(rng)-0.5);
        double a = ((2.0*M_PI)*t + dir*atan2(y,x)) * p[1];
        double r = pow(norm2sq(xy,2),p[2]);
        double sa,ca;
        sincos(a,&sa,&ca);
        *ox = ca*r;
        *oy = sa*r;
        return;
    }
    case FFR_VAR_RADIAL_BLUR: /* :1033-1043 */
    {
        double g = p[2] * rand_gaussian(rng);
        double ra = norm2(xy,2);
        double a = atan2(y,x) + p[0]*g;
        double sa,ca;
        sincos(a,&sa,&ca);
        double rz = p[1]*g - 1.0;
        *ox = ca*ra + x*rz;
        *oy = sa*ra + y*rz;
        return;
    }
    case FFR_VAR_PIE: /* :1061-1070 */
    {
        i32 sl = (i32)(rand_num(rng)*p[0] + 0.5);
        double a = p[1] + (sl + rand_num(rng)*p[2])*p[3];
        double r = rand_num(rng);
        double sa,ca;
        sincos(a,&sa,&ca);
        *ox = ca*r;
        *oy = sa*r;
        return;
    }
    case FFR_VAR_NGON: /* :1090-1100 */
    {
        double r = pow(norm2sq(xy,2),p[0]);
        double theta = atan2(y,x);
        double phi = theta - p[1]*floor(theta*p[4]);
        static const double mult[2] = {0.0,1.0};
        phi -= mult[phi > p[1]*0.5]*p[1];
        double amp = p[2]*(1.0/(cos(phi)+EPS)-1.0) + p[3];
        amp /= r + EPS;
        *ox = x*amp;
        *oy = y*amp;
        return;
    }
    case FFR_VAR_CURL: /* :1116-1126 */
    {
        double c1 = p[0], c2 = p[1];
        double re = 1.0 + c1*x + c2*(x*x - y*y);
        double im = c1*y + 2.0*c2*x*y;
        double r = 1.0 / (re*re + im*im + EPS);
        *ox = (x*re+y*im)*r;
        *oy = (y*re-x*im)*r;
        return;
    }
    case FFR_VAR_ARCH: /* :1142-1149 */
    {
        double a = p[0] * rand_num(rng) * M_PI;
        double sa,ca;
        sincos(a,&sa,&ca);
        *ox = sa;
        *oy = sa*sa/ca;
        return;
    }
    case FFR_VAR_TANGENT: /* :1159-1164 */
        *ox = sin(x)/cos(y);
        *oy = tan(y);
        return;
    case FFR_VAR_RAYS: /* :1180-1188 */
    {
        double a = p[0] * rand_num(rng) * M_PI;
        double r = p[0] / (norm2sq(xy,2) + EPS);
        double tr = tan(a) * r;
        *ox = cos(x)*tr;
        *oy = sin(y)*tr;
        return;
    }
    case FFR_VAR_BLADE: /* :1204-1210 */
    {
        double r = rand_num(rng) * p[0] * norm2(xy,2);
        double sr,cr;
        sincos(r,&sr,&cr);
        *ox = (cr+sr)*x;
        *oy = (cr-sr)*x;
        return;
    }
    case FFR_VAR_SECANT: /* :1226-1232 */
    {
        double cr = cos(p[0]*norm2(xy,2));
        double icr = 1.0/cr;
        double sign = copysign(1.0,-cr);
        *ox = x;
        *oy = icr+sign;
        return;
    }
    case FFR_VAR_TWINTRIAN: /* :1248-1258 */
    {
        double r = rand_num(rng) * p[0] * norm2(xy,2);
        double sr,cr;
        sincos(r,&sr,&cr);
        double diff = log10(sr*sr) + cr;
        if (bad_value(diff))
            diff = -30.0;
        *ox = diff*x;
        *oy = (diff-sr*M_PI)*x;
        return;
    }
    case FFR_VAR_CROSS: /* :1268-1275 */
    {
        double s = x*x - y*y;
        double r = sqrt(1.0 / (s*s + EPS));
        *ox = x*r;
        *oy = y*r;
        return;
    }
    case FFR_VAR_EXP: /* :1285-1293 */
    {
        double e = exp(x);
        double es,ec;
        sincos(y,&es,&ec);
        *ox = ec*e;
        *oy = es*e;
        return;
    }
    case FFR_VAR_LOG: /* :1303-1306 */
        *ox = log(norm2sq(xy,2));
        *oy = atan2(y,x);
        return;
    case FFR_VAR_SIN: /* :1316-1325 */
    {
        double s,c;
        sincos(x,&s,&c);
        double sh = sinh(y);
        double ch = cosh(y);
        *ox = s*ch;
        *oy = c*sh;
        return;
    }
    case FFR_VAR_COS: /* :1335-1344 */
    {
        double s,c;
        sincos(x,&s,&c);
        double ch = cosh(y);
        double sh = sinh(y);
        *ox = c*ch;
        *oy = -s*sh;
        return;
    }
    case FFR_VAR_TAN: /* :1354-1364 */
    {
        double s,c;
        sincos(2.0*x,&s,&c);
        double sh = sinh(2.0*y);
        double ch = cosh(2.0*y);
        double k = 1/(c+ch); /* Point::operator/= multiplies by 1/k, point.hpp:135-139 */
        *ox = s*k;
        *oy = sh*k;
        return;
    }
    case FFR_VAR_SEC: /* :1374-1384 */
    {
        double s,c;
        sincos(x,&s,&c);
        double sh = sinh(y);
        double ch = cosh(y);
        double k = 1/(cos(2.0*x)+cosh(2.0*y));
        *ox = (c*ch)*k;
        *oy = (s*sh)*k;
        return;
    }
    case FFR_VAR_CSC: /* :1394-1404 */
    {
        double s,c;
        sincos(x,&s,&c);
        double sh = sinh(y);
        double ch = cosh(y);
        double k = 1/(cosh(2.0*y)-cos(2.0*x));
        *ox = (s*ch)*k;
        *oy = (-c*sh)*k;
        return;
    }
    case FFR_VAR_COT: /* :1414-1424 */
    {
        double s,c;
        sincos(2.0*x,&s,&c);
        double sh = sinh(2.0*y);
        double ch = cosh(2.0*y);
        double k = 1/(ch-c);
        *ox = s*k;
        *oy = (-sh)*k;
        return;
    }
    case FFR_VAR_SINH: /* :1434-1443 */
    {
        double s,c;
        sincos(y,&s,&c);
        double sh = sinh(x);
        double ch = cosh(x);
        *ox = sh*c;
        *oy = ch*s;
        return;
    }
    case FFR_VAR_COSH: /* :1453-1462 */
    {
        double s,c;
        sincos(y,&s,&c);
        double sh = sinh(x);
        double ch = cosh(x);
        *ox = ch*c;
        *oy = sh*s;
        return;
    }
    case FFR_VAR_TANH: /* :1472-1482 */
    {
        double s,c;
        sincos(2.0*y,&s,&c);
        double sh = sinh(2.0*x);
        double ch = cosh(2.0*x);
        double k = 1/(c+ch);
        *ox = sh*k;
        *oy = s*k;
        return;
    }
    case FFR_VAR_SECH: /* :1492-1502 */
    {
        double s,c;
        sincos(y,&s,&c);
        double sh = sinh(x);
        double ch = cosh(x);
        double k = 1/(cos(2.0*y)+cosh(2.0*x));
        *ox = (c*ch)*k;
        *oy = (-s*sh)*k;
        return;
    }
    case FFR_VAR_CSCH: /* :1512-1522 */
    {
        double s,c;
        sincos(y,&s,&c);
        double sh = sinh(x);
        double ch = cosh(x);
        double k = 1/(cosh(2.0*x)-cos(2.0*y));
        *ox = (sh*c)*k;
        *oy = (-ch*s)*k;
        return;
    }
    case FFR_VAR_COTH: /* :1532-1542 */
    {
        double s,c;
        sincos(2.0*y,&s,&c);
        double sh = sinh(2.0*x);
        double ch = cosh(2.0*x);
        double k = 1/(ch-c);
        *ox = sh*k;
        *oy = s*k;
        return;
    }
    case FFR_VAR_AUGER: /* :1560-1569 */
    {
        double s = sin(p[0]*x);
        double t = sin(p[0]*y);
        double dy = y + p[1]*(p[2] + fabs(y))*s;
        double dx = x + p[1]*(p[2] + fabs(x))*t;
        *ox = x+p[3]*(dx-x);
        *oy = dy;
        return;
    }
    case FFR_VAR_FLUX: /* :1586-1598 (sincosg outputs named the other way round) */
    {
        double xpw = x + p[1];
        double xmw = x - p[1];
        double y2 = y*y;
        double avgr = p[0] * sqrt(sqrt(y2+xpw*xpw)/sqrt(y2+xmw*xmw));
        double avga = (atan2(y,xmw) - atan2(y,xpw)) * 0.5;
        double c,s;
        sincos(avga,&c,&s); /* c = sin, s = cos, as written in the reference */
        *ox = c*avgr;
        *oy = s*avgr;
        return;
    }
    case FFR_VAR_MOBIUS: /* :1616-1627 */
    {
        double re_u = p[0]*x - p[1]*y + p[2];
        double im_u = p[0]*y + p[1]*x + p[3];
        double re_v = p[4]*x - p[5]*y + p[6];
        double im_v = p[4]*y + p[5]*x + p[7];
        double rad = 1.0 / (re_v*re_v + im_v*im_v + EPS);
        *ox = (re_u*re_v+im_u*im_v)*rad;
        *oy = (im_u*re_v-re_u*im_v)*rad;
        return;
    }
    case FFR_VAR_SCRY: /* :1643-1648 */
    {
        double t = norm2sq(xy,2);
        double r = 1.0 / (sqrt(t) * (t + 1.0/(p[0] + EPS)));
        *ox = x*r;
        *oy = y*r;
        return;
    }
    case FFR_VAR_SPLIT: /* :1664-1671 */
    {
        double xs = copysign(1.0,cos(x*p[0]));
        double ys = copysign(1.0,cos(y*p[1]));
        *ox = x*ys;
        *oy = y*xs;
        return;
    }
    case FFR_VAR_STRIPES: /* :1687-1694 */
    {
        double rx = floor(x + 0.5);
        double ox_ = x - rx;
        *ox = ox_*p[0]+rx;
        *oy = y+ox_*ox_*p[1];
        return;
    }
    case FFR_VAR_WEDGE: /* :1713-1722 */
    {
        double r = norm2(xy,2);
        double a = atan2(y,x) + p[0]*r;
        double c = floor((p[1]*a + M_PI) * (M_1_PI*0.5));
        a = a*p[4] + c*p[2];
        double sa,ca;
        sincos(a,&sa,&ca);
        double k = r+p[3];
        *ox = ca*k;
        *oy = sa*k;
        return;
    }
    case FFR_VAR_WEDGE_JULIA: /* :1744-1754 */
    {
        double r = pow(norm2sq(xy,2),p[0]);
        i32 tr = (i32)(p[1] * rand_num(rng));
        double a = (atan2(y,x) + (2.0*M_PI)*tr) * p[2];
        double c = floor((p[3]*a + M_PI) * (M_1_PI*0.5));
        double sa,ca;
        a = a*p[5] + c*p[4];
        sincos(a,&sa,&ca);
        *ox = ca*r;
        *oy = sa*r;
        return;
    }
    case FFR_VAR_WEDGE_SPH: /* :1773-1782 */
    {
        double r = 1.0 / (norm2(xy,2) + EPS);
        double a = atan2(y,x) + p[0]*r;
        double c = floor((p[1]*a + M_PI) * (M_1_PI*0.5));
        double sa,ca;
        a = a*p[2] + c*p[3];
        sincos(a,&sa,&ca);
        double k = r+p[4];
        *ox = ca*k;
        *oy = sa*k;
        return;
    }
    case FFR_VAR_WHORL: /* :1800-1808 */
    {
        double r = norm2(xy,2);
        double a = atan2(y,x);
        a += p[r >= p[2]] / (p[2] - r);
        double sa,ca;
        sincos(a,&sa,&ca);
        *ox = ca*r;
        *oy = sa*r;
        return;
    }
    case FFR_VAR_SUPERSHAPE: /* :1829-1840 */
    {
        double theta = p[0]*atan2(y,x) + M_PI_4;
        double st,ct;
        sincos(theta,&st,&ct);
        double t1 = pow(fabs(ct),p[2]);
        double t2 = pow(fabs(st),p[3]);
        double tr = norm2(xy,2);
        double r = (p[4]*rand_num(rng) + (1.0-p[4])*tr) - p[5];
        r *= pow(t1+t2,p[1]) / tr;
        *ox = x*r;
        *oy = y*r;
        return;
    }
    case FFR_VAR_FLOWER: /* :1856-1862 */
    {
        double theta = atan2(y,x);
        double r = (rand_num(rng) - p[1]) * cos(p[0]*theta);
        r /= norm2(xy,2) + EPS;
        *ox = x*r;
        *oy = y*r;
        return;
    }
    case FFR_VAR_CONIC: /* :1878-1884 */
    {
        double tr = norm2(xy,2);
        double ct = x / (tr + EPS);
        double r = (rand_num(rng) - p[1]) * p[0] / (tr + tr*p[0]*ct);
        *ox = x*r;
        *oy = y*r;
        return;
    }
    case FFR_VAR_PARABOLA: /* :1900-1907 */
    {
        double sr,cr;
        sincos(norm2(xy,2),&sr,&cr);
        double px = p[0]*sr*sr*rand_num(rng);
        double py = p[1]*cr*rand_num(rng);
        *ox = px;
        *oy = py;
        return;
    }
    case FFR_VAR_BIPOLAR: /* :1922-1932 */
    {
        double x2y2 = norm2sq(xy,2);
        double t = x2y2 + 1.0;
        double x2 = 2.0*x;
        double yy = 0.5*atan2(2.0*y,x2y2-1.0) + p[0];
        yy -= M_PI * floor(yy*M_1_PI + 0.5);
        *ox = log((t+x2)/(t-x2));
        *oy = yy;
        return;
    }
    case FFR_VAR_BOARDERS: /* :1951-1981 */
    {
        double rx = rint(x);
        double ry = rint(y);
        double ox_ = x - rx;
        double oy_ = y - ry;
        if (rand_num(rng) >= p[0])
        {
            *ox = ox_*0.5+rx;
            *oy = oy_*0.5+ry;
        }
        else
        {
            double mag = 1.0 - p[0];
            if (fabs(ox_) >= fabs(oy_))
            {
                double s = copysign(mag,ox_);
                *ox = ox_*0.5 + rx + s;
                *oy = oy_*0.5 + ry + s*oy_/ox_;
            }
            else
            {
                double s = copysign(mag,oy_);
                *ox = ox_*0.5 + rx + s*ox_/oy_;
                *oy = oy_*0.5 + ry + s;
            }
        }
        return;
    }
    case FFR_VAR_BUTTERFLY: /* :1994-2001 */
    {
        double y2 = 2.0*y;
        double r = sqrt(fabs(x*y) / (x*x + y2*y2 + EPS));
        *ox = x*r;
        *oy = y2*r;
        return;
    }
    case FFR_VAR_CELL: /* :2017-2031 */
    {
        double size = p[0], invsize = p[1];
        double cx = floor(x * invsize);
        double cy = floor(y * invsize);
        double dx = x - cx*size;
        double dy = y - cy*size;
        double xs = copysign(2.0,cx);
        double ys = copysign(2.0,cy);
        double x2 = cx * xs;
        double y2 = cy * ys;
        x2 -= (double)(cx < 0);
        y2 -= (double)(cy < 0);
        *ox = dx+x2*size;
        *oy = -dy-y2*size;
        return;
    }
    case FFR_VAR_CPOW: /* :2051-2059 */
    {
        double a = atan2(y,x);
        double lnr = 0.5 * log(norm2sq(xy,2));
        double ang = p[1]*a + p[2]*lnr + p[0]*floor(p[3]*rand_num(rng));
        double sa,ca;
        sincos(ang,&sa,&ca);
        double e = exp(p[1]*lnr - p[2]*a);
        *ox = ca*e;
        *oy = sa*e;
        return;
    }
    case FFR_VAR_CURVE: /* :2082-2089 */
    {
        double vx = p[2]*exp(-y*y*p[0]);
        double vy = p[3]*exp(-x*x*p[1]);
        *ox = x + vx;
        *oy = y + vy;
        return;
    }
    case FFR_VAR_EDISC: /* :2103-2118 */
    {
        double tmp = norm2sq(xy,2) + 1.0;
        double tmp2 = 2.0*x;
        double xmax = 0.5*(sqrt(tmp+tmp2) + sqrt(tmp-tmp2));
        double a1 = log(xmax + sqrt(xmax-1.0));
        double a2 = -acos(x/xmax);
        double s1,c1;
        sincos(a1,&s1,&c1);
        double s2 = sinh(a2);
        double c2 = cosh(a2);
        s1 *= copysign(1.0,-y);
        *ox = c2*c1;
        *oy = s2*s1;
        return;
    }
    case FFR_VAR_ELLIPTIC: /* :2128-2142 */
    {
        double tmp = norm2sq(xy,2) + 1.0;
        double x2 = 2.0*x;
        double xmax = 0.5*(sqrt(tmp+x2) + sqrt(tmp-x2));
        double a = x/xmax;
        double b = 1.0 - a*a;
        double ssx = xmax - 1.0;
        b = b < 0.0 ? 0.0 : sqrt(b);
        ssx = ssx < 0.0 ? 0.0 : sqrt(ssx);
        *ox = atan2(a,b);
        *oy = copysign(1.0,y)*log(xmax+ssx);
        return;
    }
    case FFR_VAR_ESCHER: /* :2161-2169 */
    {
        double a = atan2(y,x);
        double lnr = 0.5*log(norm2sq(xy,2));
        double n = p[0]*a + p[1]*lnr;
        double sn,cn;
        sincos(n,&sn,&cn);
        double e = exp(p[0]*lnr - p[1]*a);
        *ox = cn*e;
        *oy = sn*e;
        return;
    }
    case FFR_VAR_FOCI: /* :2179-2189 */
    {
        double expx = 0.5*exp(x);
        double expnx = 0.25/expx;
        double sn,cn;
        sincos(y,&sn,&cn);
        double tmp = 1.0 / (expx + expnx - cn);
        *ox = (expx-expnx)*tmp;
        *oy = sn*tmp;
        return;
    }
    case FFR_VAR_LAZYSUSAN: /* :2210-2227 */
    {
        double lx = x - p[0];
        double ly = y + p[1];
        double r = hypot(lx,ly);
        if (r < p[5])
        {
            double a = atan2(ly,lx) + p[2] + p[3]*(p[5] - r);
            double sa,ca;
            sincos(a,&sa,&ca);
            *ox = r*ca+p[0];
            *oy = r*sa-p[1];
        }
        else
        {
            r = 1.0 + p[4] / (r + EPS);
            *ox = r*lx+p[0];
            *oy = r*ly-p[1];
        }
        return;
    }
    case FFR_VAR_LOONIE: /* :2244-2252 */
    {
        double r2 = norm2sq(xy,2);
        double w2 = p[1];
        double r = p[0];
        if (r2 < w2) r *= sqrt(w2/(r2 + EPS) - 1.0);
        *ox = x*r;
        *oy = y*r;
        return;
    }
    case FFR_VAR_OSCOPE: /* :2271-2279 */
    {
        double damp = exp(-fabs(x)*p[2]);
        double t = p[1] * damp * cos(p[0]*x) + p[3];
        double yy = copysign(1.0,fabs(y)-t) * y;
        *ox = x;
        *oy = yy;
        return;
    }
    case FFR_VAR_POPCORN: /* :2296-2301 */
    {
        double dx = p[0]*sin(tan(y*p[2]));
        double dy = p[1]*sin(tan(x*p[2]));
        *ox = x + dx;
        *oy = y + dy;
        return;
    }
    default:
        *ox = *oy = NAN;
        return;
    }
}

/* Variation::calc for one variation; t in, o out (D entries) */
static void var_calc(const ffr_variation *v, isaac64 *rng, int D, const double *t,
        double *o)
{
    const double *p = v->params;
    int i;
    if (v->op >= FFR_VAR_FIRST_2D && v->op <= FFR_VAR_LAST_2D)
    {
        /* VariationFrom2D::calc_h :94-105 */
        double ox,oy;
        if (D == 2)
        {
            calc2d(v,rng,t[0],t[1],&ox,&oy);
            o[0] = ox;
            o[1] = oy;
        }
        else
        {
            calc2d(v,rng,t[v->axis_x],t[v->axis_y],&ox,&oy);
            for (i = 0; i < D; ++i)
                o[i] = 0;
            o[v->axis_x] = ox;
            o[v->axis_y] = oy;
        }
        return;
    }
    switch (v->op)
    {
    case FFR_VAR_LINEAR: /* :173-176 */
        for (i = 0; i < D; ++i)
            o[i] = t[i];
        return;
    case FFR_VAR_SINUSOIDAL: /* :187-190 */
        for (i = 0; i < D; ++i)
            o[i] = sin(t[i]);
        return;
    case FFR_VAR_SPHERICAL: /* :201-207 */
    {
        double r = 1.0 / (norm2sq(t,D) + EPS);
        for (i = 0; i < D; ++i)
            o[i] = t[i]*r;
        return;
    }
    case FFR_VAR_BENT: /* :227-238 */
        for (i = 0; i < D; ++i)
        {
            double x = t[i];
            if (x < 0.0)
                x *= p[i];
            else
                x *= p[4+i];
            o[i] = x;
        }
        return;
    case FFR_VAR_RECTANGLES: /* :253-266 */
        for (i = 0; i < D; ++i)
        {
            double q = p[i];
            double x = t[i];
            if (q == 0.0)
                o[i] = x;
            else
                o[i] = (2.0*floor(x/q) + 1.0)*q - x;
        }
        return;
    case FFR_VAR_FISHEYE: /* :285-290 */
    {
        double r = 1.0 / (norm2(t,D) + p[0]);
        for (i = 0; i < D; ++i)
            o[i] = t[i]*r;
        return;
    }
    case FFR_VAR_BUBBLE: /* :307-312 */
    {
        double r = 1.0 / (norm2sq(t,D) + p[0]);
        for (i = 0; i < D; ++i)
            o[i] = t[i]*r;
        return;
    }
    case FFR_VAR_NOISE: /* :322-328 */
    {
        double r = rand_num(rng);
        double dir[3];
        rand_direction(rng,D,dir);
        for (i = 0; i < D; ++i)
            o[i] = (t[i]*dir[i])*r;
        return;
    }
    case FFR_VAR_BLUR: /* :338-345 */
    {
        double r = rand_num(rng);
        double dir[3];
        rand_direction(rng,D,dir);
        for (i = 0; i < D; ++i)
            o[i] = dir[i]*r;
        return;
    }
    case FFR_VAR_GAUSSIAN_BLUR: /* :355-362 */
    case FFR_VAR_PRE_BLUR:      /* :437-444 */
    {
        double r = rand_gaussian(rng);
        double dir[3];
        rand_direction(rng,D,dir);
        for (i = 0; i < D; ++i)
            o[i] = dir[i]*r;
        return;
    }
    case FFR_VAR_SQUARE_NOISE: /* :372-376, randPoint2 flame_rng.hpp:161-168 */
        for (i = 0; i < D; ++i)
            o[i] = rand_num(rng) - 0.5;
        return;
    case FFR_VAR_SEPARATION: /* :394-403 */
        for (i = 0; i < D; ++i)
        {
            double s = copysign(1.0,t[i]);
            o[i] = s * (sqrt(t[i]*t[i] + p[i]) - s*p[4+i]);
        }
        return;
    case FFR_VAR_SPLITS: /* :418-427 */
        for (i = 0; i < D; ++i)
        {
            double s = copysign(1.0,t[i]);
            o[i] = t[i] + s*p[i];
        }
        return;
    case FFR_VAR_MODULUS: /* :461-468 */
        for (i = 0; i < D; ++i)
            o[i] = t[i] - p[i]*floor(t[i]*p[4+i] + 0.5);
        return;
    case FFR_VAR_CELLN: /* :486-499 */
        for (i = 0; i < D; ++i)
        {
            double x = floor(t[i] * p[4+i]);
            double dx = t[i] - x*p[i];
            double xs = copysign(2.0,x);
            double x2 = x * xs;
            x2 -= (double)(x < 0);
            o[i] = dx + x2*p[i];
        }
        return;
    case FFR_VAR_SPHERICAL_P: /* :2322-2326 */
    {
        double r = 1.0 / (normsum_p(t,D,p[0]) + EPS);
        for (i = 0; i < D; ++i)
            o[i] = t[i]*r;
        return;
    }
    case FFR_VAR_UNIT_SPHERE: /* :2336-2340 */
    {
        double r = 1.0 / (norm2(t,D) + EPS);
        for (i = 0; i < D; ++i)
            o[i] = t[i]*r;
        return;
    }
    case FFR_VAR_UNIT_SPHERE_P: /* :2357-2361, norm(T p) point.hpp:291-294 */
    {
        double r = 1.0 / (pow(normsum_p(t,D,p[0]),1.0/p[0]) + EPS);
        for (i = 0; i < D; ++i)
            o[i] = t[i]*r;
        return;
    }
    case FFR_VAR_UNIT_CUBE: /* :2371-2375 */
    {
        double r = 1.0 / (norminf(t,D) + EPS);
        for (i = 0; i < D; ++i)
            o[i] = t[i]*r;
        return;
    }
    default:
        for (i = 0; i < D; ++i)
            o[i] = NAN;
        return;
    }
}

/* Affine::apply_to, types/affine.hpp:104-110 with the dot product of point.hpp:228-234:
   ret[i] = b[i] + (((0 + A[i][0]*x[0]) + A[i][1]*x[1]) ...) */
static inline void affine_apply(int D, const double *A, const double *b, const double *x,
        double *out)
{
    for (int i = 0; i < D; ++i)
    {
        double dot = 0;
        for (int j = 0; j < D; ++j)
            dot += A[i*D+j] * x[j];
        out[i] = b[i] + dot;
    }
}

/* XForm::applyIteration, types/xform.hpp:211-227 */
static void xform_apply(const ffr_xform *xf, isaac64 *rng, int D, const double *p,
        double *out)
{
    double t[3],v[3],c[3];
    int i;
    uint32_t k;
    if (D < 3 || xf->has_pre)
        affine_apply(D,xf->pre_A,xf->pre_b,p,t);
    else
        for (i = 0; i < D; ++i) t[i] = p[i];
    for (i = 0; i < D; ++i)
        v[i] = 0;
    for (k = 0; k < xf->num_vars; ++k)
    {
        var_calc(&xf->vars[k],rng,D,t,c);
        /* v += weight * calc(t): operator*(T,Point) -> calc[i]*weight (point.hpp:215-225) */
        for (i = 0; i < D; ++i)
            v[i] += c[i] * xf->vars[k].weight;
    }
    if (D < 3 || xf->has_post)
        affine_apply(D,xf->post_A,xf->post_b,v,out);
    else
        for (i = 0; i < D; ++i) out[i] = v[i];
}

/* Flame::getRandomXForm, types/flame.hpp:212-219 */
static inline uint32_t select_xform(const ffr_flame_desc *fl, isaac64 *rng)
{
    uint32_t i = 0;
    double r = rand_num(rng);
    while (fl->xfcw[i] < r)
        ++i;
    return i;
}

/* one chain's state: RenderIterator, renderers/render_iterator.hpp:38-49 */
typedef struct
{
    double p[3], pf[3];
    double c[FFR_MAX_COLOR_DIMS], cf[FFR_MAX_COLOR_DIMS];
} chain_state;

/* RenderIterator::_init, render_iterator.hpp:52-60 */
static void chain_init(const ffr_flame_desc *fl, isaac64 *rng, chain_state *st)
{
    int D = (int)fl->dims;
    double q[3];
    for (int i = 0; i < D; ++i)
        st->p[i] = 2.0*rand_num(rng) - 1.0; /* randPoint, flame_rng.hpp:151-158 */
    for (int s = 0; s < SETTLE_ITERS; ++s)
    {
        uint32_t xi = select_xform(fl,rng);
        xform_apply(&fl->xforms[xi],rng,D,st->p,q);
        for (int i = 0; i < D; ++i)
            st->p[i] = q[i];
    }
    for (uint32_t i = 0; i < fl->color_dims; ++i)
        st->c[i] = rand_num(rng);
}

/* RenderIterator::iterate, render_iterator.hpp:106-139; returns the xform's JSON id */
static u64 chain_iterate(const ffr_flame_desc *fl, isaac64 *rng, chain_state *st)
{
    int D = (int)fl->dims;
    uint32_t r = fl->color_dims;
    double q[3];
    uint32_t xi = select_xform(fl,rng);
    const ffr_xform *xf = &fl->xforms[xi];
    double s = xf->color_speed;
    xform_apply(xf,rng,D,st->p,q);
    for (int i = 0; i < D; ++i)
        st->p[i] = q[i];
    if (xf->has_color)
        for (uint32_t i = 0; i < r; ++i)
            st->c[i] = (1.0-s)*st->c[i] + s*xf->color[i];
    if (fl->has_final)
    {
        const ffr_xform *xff = fl->final_xform;
        s = xff->color_speed;
        xform_apply(xff,rng,D,st->p,st->pf);
        if (xff->has_color)
            for (uint32_t i = 0; i < r; ++i)
                st->cf[i] = (1.0-s)*st->c[i] + s*xff->color[i];
        else
            for (uint32_t i = 0; i < r; ++i)
                st->cf[i] = st->c[i];
    }
    else
    {
        for (int i = 0; i < D; ++i)
            st->pf[i] = st->p[i];
        for (uint32_t i = 0; i < r; ++i)
            st->cf[i] = st->c[i];
    }
    return xf->id;
}

/* NaN pf. In the reference a NaN coordinate passes _in_bounds (all comparisons
   false, render_iterator.hpp:72-79) and then hits an undefined double->size_t cast
   (buffer_renderer.hpp:202). Default here, and on the device: NaN is out of bounds.
   With this flag set the oracle instead reproduces what the x86-64 reference binary
   happens to do (cvttsd2si gives 2^63, which vanishes from the byte offset mod 2^64, so the
   NaN coordinate contributes 0 to the index and the sample IS counted) -- used only to pin
   this file against oracle/_ref on flames that produce NaN. */
static int g_emulate_x86_nan_cast = 0;
void oracle_set_nan_emulation(int on) { g_emulate_x86_nan_cast = on; }

typedef struct
{
    const ffr_flame_desc *fl;
    double mult_d[3];
    u64 mult_i[3];
    u64 cell;            /* 1 + color_dims */
    u64 *buffer;         /* cells * cell elements, u64 / double interleaved */
    int atomic;          /* shared buffer written by several threads */
} render_target;

static inline void add_double(u64 *slot, double v, int atomic)
{
    if (!atomic)
    {
        double cur;
        memcpy(&cur,slot,8);
        cur += v;
        memcpy(slot,&cur,8);
    }
    else
    {
        u64 old = __atomic_load_n(slot,__ATOMIC_RELAXED), neu;
        double cur;
        do
        {
            memcpy(&cur,&old,8);
            cur += v;
            memcpy(&neu,&cur,8);
        }
        while (!__atomic_compare_exchange_n(slot,&old,neu,1,__ATOMIC_RELAXED,__ATOMIC_RELAXED));
    }
}

/* BufferRenderer::_render_batch, renderers/buffer_renderer.hpp:150-250: one chain.
   Returns 0 when all samples were iterated, 1 when the bad value limit stopped it. */
static int render_chain(render_target *tg, u64 seed, u64 samples, u64 bv_limit,
        ffr_stats *st, u64 *n_bad_chain)
{
    const ffr_flame_desc *fl = tg->fl;
    int D = (int)fl->dims;
    uint32_t r = fl->color_dims;
    isaac64 rng;
    chain_state cs;
    u64 bad = 0;
    isaac_seed(&rng,seed);           /* rng::setSeed((u64)seed_k) */
    chain_init(fl,&rng,&cs);         /* RenderIterator ctor :155 */
    for (; samples > 0; --samples)
    {
        u64 xf_id = chain_iterate(fl,&rng,&cs);
        ++st->s_iter;
        ++st->xf_dist[xf_id];
        int isbad = 0;
        for (int i = 0; i < D; ++i)
            if (bad_value(cs.p[i]))
                isbad = 1;
        if (isbad) /* :175-186 */
        {
            if (st->n_bad < FFR_MAX_BAD_RECORDED)
            {
                st->bad_xf[st->n_bad] = xf_id;
                for (int i = 0; i < D; ++i)
                    st->bad_pt[st->n_bad][i] = cs.p[i];
            }
            ++st->n_bad;
            ++bad;
            /* renderSeeded clears the lists per call (:352-353), so the limit is per chain */
            if (bad > bv_limit)
                break;
            chain_init(fl,&rng,&cs); /* pf, cf keep the stale values (Q3) */
        }
        for (int i = 0; i < D; ++i) /* :188-194, on the (possibly re-initialised) p */
        {
            if (cs.p[i] < st->pt_min[i])
                st->pt_min[i] = cs.p[i];
            if (cs.p[i] > st->pt_max[i])
                st->pt_max[i] = cs.p[i];
        }
        /* _in_bounds on pf, render_iterator.hpp:72-79. A NaN pf passes this test in the
           reference and then hits an undefined double->size_t cast; the fenced behaviour
           (SURVEY Q4), identical on the device, is: NaN is out of bounds. */
        int inb = 1;
        for (int i = 0; i < D; ++i)
        {
            if (g_emulate_x86_nan_cast)
            {
                if (cs.pf[i] < fl->bounds_lo[i] || cs.pf[i] > fl->bounds_hi[i])
                    inb = 0;
            }
            else if (!(cs.pf[i] >= fl->bounds_lo[i] && cs.pf[i] <= fl->bounds_hi[i]))
                inb = 0;
        }
        if (!inb)
            continue;
        ++st->s_plot;
        /* :202-209 */
        u64 bi = 0;
        for (int i = 0; i < D; ++i)
        {
            double sc = (cs.pf[i] - fl->bounds_lo[i]) * tg->mult_d[i];
            u64 di = isnan(sc) ? 0 : (u64)sc; /* NaN only reachable with the emulation flag */
            bi += di * tg->mult_i[i];
        }
        u64 *bptr = tg->buffer + bi*tg->cell;
        if (tg->atomic)
            __atomic_fetch_add(bptr,1,__ATOMIC_RELAXED);
        else
            ++*bptr;
        for (uint32_t i = 0; i < r; ++i)
            add_double(bptr+1+i,cs.cf[i],tg->atomic);
    }
    *n_bad_chain = bad;
    return samples == 0 ? 0 : 1;
}

static u64 splitmix64(u64 x)
{
    u64 z = x + 0x9e3779b97f4a7c15uLL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9uLL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebuLL;
    return z ^ (z >> 31);
}

static int setup_target(render_target *tg, const ffr_flame_desc *fl, void *buffer, int atomic)
{
    /* BufferRenderer::_init, buffer_renderer.hpp:114-140 */
    const double scale_adjust_down = 1.0 - (double)(float)(1.0 / (double)(1L << 52));
    u64 cells = 1;
    tg->fl = fl;
    for (uint32_t i = 0; i < fl->dims; ++i)
    {
        tg->mult_d[i] = (double)(fl->size[i]) / (fl->bounds_hi[i] - fl->bounds_lo[i]);
        tg->mult_d[i] *= scale_adjust_down;
        tg->mult_i[i] = cells;
        cells *= fl->size[i];
        if (cells >= (1uLL << 48))
            return -1;
    }
    tg->cell = 1 + fl->color_dims;
    tg->buffer = (u64*)buffer;
    tg->atomic = atomic;
    return 0;
}

static void stats_init(ffr_stats *st)
{
    memset(st,0,sizeof(*st));
    for (int i = 0; i < 3; ++i)
    {
        st->pt_min[i] = INFINITY;
        st->pt_max[i] = -INFINITY;
    }
}

static void stats_merge(ffr_stats *dst, const ffr_stats *src, int D)
{
    dst->s_iter += src->s_iter;
    dst->s_plot += src->s_plot;
    for (int i = 0; i < FFR_MAX_XFORMS; ++i)
        dst->xf_dist[i] += src->xf_dist[i];
    for (int i = 0; i < D; ++i)
    {
        if (src->pt_min[i] < dst->pt_min[i]) dst->pt_min[i] = src->pt_min[i];
        if (src->pt_max[i] > dst->pt_max[i]) dst->pt_max[i] = src->pt_max[i];
    }
    for (u64 k = 0; k < src->n_bad && k < FFR_MAX_BAD_RECORDED; ++k)
    {
        if (dst->n_bad + k < FFR_MAX_BAD_RECORDED)
        {
            dst->bad_xf[dst->n_bad + k] = src->bad_xf[k];
            memcpy(dst->bad_pt[dst->n_bad + k],src->bad_pt[k],sizeof(src->bad_pt[k]));
        }
    }
    dst->n_bad += src->n_bad;
}

typedef struct
{
    render_target tg;
    u64 base_seed, chain_first, chain_count, chain_len, last_len, bv_limit;
    u64 tid, nthreads;
    ffr_stats *st;
    int ret;
} worker_arg;

static void *worker(void *argp)
{
    worker_arg *a = (worker_arg*)argp;
    a->ret = 0;
    for (u64 k = a->tid; k < a->chain_count; k += a->nthreads)
    {
        u64 len = (k+1 == a->chain_count && a->last_len) ? a->last_len : a->chain_len;
        u64 nb;
        if (render_chain(&a->tg,splitmix64(a->base_seed + a->chain_first + k),len,
                a->bv_limit,a->st,&nb))
        {
            a->ret = 1;
            break;
        }
    }
    return NULL;
}

/* ---------------- exported (ctypes) ---------------- */

u64 oracle_splitmix64(u64 x) { return splitmix64(x); }

void oracle_isaac_words(u64 seed, u64 n, u64 *out)
{
    isaac64 s;
    isaac_seed(&s,seed);
    for (u64 i = 0; i < n; ++i)
        out[i] = isaac_next(&s);
}

void oracle_rand_nums(u64 seed, u64 n, double *out)
{
    isaac64 s;
    isaac_seed(&s,seed);
    for (u64 i = 0; i < n; ++i)
        out[i] = rand_num(&s);
}

/* Sum over chains k in [chain_first, chain_first+chain_count) of
   { rng::setSeed(splitmix64(base_seed+k)); renderSeeded(len_k,len_k,bv_limit); } ADDED into
   buffer (which the caller zeroes or pre-loads, like -i). nthreads > 1 spreads chains over
   threads with atomic adds: counts stay exact, colour sums change in the last bits.
   stats is overwritten. Returns 0, or 1 if a chain stopped on the bad value limit. */
int oracle_render_chains(const ffr_flame_desc *fl, u64 base_seed, u64 chain_first,
        u64 chain_count, u64 chain_len, u64 last_len, u64 bv_limit, void *buffer,
        ffr_stats *stats, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if ((u64)nthreads > chain_count) nthreads = chain_count ? (int)chain_count : 1;
    worker_arg *args = (worker_arg*)calloc(nthreads,sizeof(worker_arg));
    ffr_stats *sts = (ffr_stats*)calloc(nthreads,sizeof(ffr_stats));
    pthread_t *th = (pthread_t*)calloc(nthreads,sizeof(pthread_t));
    int ret = 0;
    for (int t = 0; t < nthreads; ++t)
    {
        if (setup_target(&args[t].tg,fl,buffer,nthreads > 1))
        {
            ret = -1;
            goto done;
        }
        stats_init(&sts[t]);
        args[t].base_seed = base_seed;
        args[t].chain_first = chain_first;
        args[t].chain_count = chain_count;
        args[t].chain_len = chain_len;
        args[t].last_len = last_len;
        args[t].bv_limit = bv_limit;
        args[t].tid = t;
        args[t].nthreads = nthreads;
        args[t].st = &sts[t];
    }
    if (nthreads == 1)
        worker(&args[0]);
    else
    {
        for (int t = 0; t < nthreads; ++t)
            pthread_create(&th[t],NULL,worker,&args[t]);
        for (int t = 0; t < nthreads; ++t)
            pthread_join(th[t],NULL);
    }
    stats_init(stats);
    for (int t = 0; t < nthreads; ++t)
    {
        stats_merge(stats,&sts[t],(int)fl->dims);
        if (args[t].ret)
            ret = 1;
    }
done:
    free(args);
    free(sts);
    free(th);
    return ret;
}

/* one application of xform #xf_index (-1 = final) to each point with its own seeded rng:
   XForm::applyIteration, xform.hpp:211-227 */
int oracle_iterate_points(const ffr_flame_desc *fl, int64_t xf_index, u64 n, const u64 *seeds,
        const double *pts_in, double *pts_out)
{
    const ffr_xform *xf;
    int D = (int)fl->dims;
    if (xf_index < 0)
    {
        if (!fl->has_final)
            return -1;
        xf = fl->final_xform;
    }
    else
    {
        if ((u64)xf_index >= fl->num_xforms)
            return -1;
        xf = &fl->xforms[xf_index];
    }
    for (u64 i = 0; i < n; ++i)
    {
        isaac64 rng;
        isaac_seed(&rng,seeds[i]);
        xform_apply(xf,&rng,D,pts_in + D*i,pts_out + D*i);
    }
    return 0;
}

/* ---------------- tone map ----------------
   Restates render_image (src/ffr_img.cpp:199-309) + ImageRenderer::getValueBounds /
   renderGrayImage / renderColorImageRGB (renderers/image_renderer.hpp:112-192) for one pixel
   format. PINNED: ffr-img cannot be built here as a program (Boost and libpng are absent),
   but its pixel arithmetic is compiled from the reference's own sources by
   oracle/ref_img_harness.cpp (`make -C oracle refimg`); this function is bit-identical to it
   for every mode x bit depth x gamma (tests/golden/golden_img.json, tests/test_golden.py,
   tests/test_oracle_vs_reference.py). mode: 1 mono, 2 gray, 3 rgb; bits 8 or 16; out is
   u8 or u16 samples, channels interleaved. Returns 0, or -1 "histogram is (probably) empty".
   The double -> pixel casts of NaN (rgb, count 0: 0/0) and of 2^bits are undefined in the
   reference; as on the device they yield 0 and the top code. */
int oracle_tonemap(const u64 *buf, u64 cells, uint32_t cellsz, int mode, int bits, double gamma,
        void *out, u64 *hist_min, u64 *hist_max, double *sc_min, double *sc_max)
{
    /* pix_scale_v<pix_t,double>, constants.hpp:77-91 */
    const double scale_adjust_down = 1.0 - (double)(float)(1.0 / (double)(1L << 52));
    const double pix_scale = (bits == 8 ? 256.0 : 65536.0) * scale_adjust_down;
    const double top = bits == 8 ? 255.0 : 65535.0;
    u64 minh = buf[0], maxh = buf[0];
    double mn = log(1 + (double)buf[0]), mx = mn;
    for (u64 i = 0; i < cells; ++i) /* getValueBounds :116-126 with funch and func1 */
    {
        u64 n = buf[i*cellsz];
        double l = log(1 + (double)n);
        if (n < minh) minh = n;
        if (n > maxh) maxh = n;
        if (l < mn) mn = l;
        if (l > mx) mx = l;
    }
    if (hist_min) *hist_min = minh;
    if (hist_max) *hist_max = maxh;
    if (sc_min) *sc_min = mn;
    if (sc_max) *sc_max = mx;
    if (mx < EPS) /* :231-232 */
        return -1;
    double gp = 1.0 / gamma; /* :235 */
    for (u64 i = 0; i < cells; ++i)
    {
        const u64 *cell = buf + i*cellsz;
        u64 n = cell[0];
        double vals[3];
        int ch = 1;
        if (mode == 1) /* :259-263 */
            vals[0] = n != 0 ? 1.0 : 0.0;
        else
        {
            double l = log(1 + (double)n) / mx; /* :240-241 */
            double ll = pow(l,gp);
            if (mode == 2)
                vals[0] = ll;
            else /* :284-294 */
            {
                double h = (double)n;
                ch = 3;
                for (int c = 0; c < 3; ++c)
                {
                    double col;
                    memcpy(&col,&cell[1+c],8);
                    vals[c] = ll*(col / h);
                }
            }
        }
        for (int c = 0; c < ch; ++c)
        {
            double v = vals[c] * pix_scale; /* image_renderer.hpp:162,185-187 */
            if (isnan(v) || v < 0.0) v = 0.0;
            if (v > top) v = top;
            if (bits == 8 || mode == 1)
                ((uint8_t*)out)[i*ch+c] = (uint8_t)v;
            else
                ((uint16_t*)out)[i*ch+c] = (uint16_t)v;
        }
    }
    return 0;
}
