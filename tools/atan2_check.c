#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
// CPU check of the fast atan2 used by k_gradient (sift_pyocl_b200/csrc/common.cuh: cr_atan2f_fast) against glibc:
// gcc -O2 -mfma -ffp-contract=off -o tools/atan2_check tools/atan2_check.c -lm && tools/atan2_check 200000000 [t-perturbation, e.g. 1.000001] [seed-perturbation, e.g. 1.0000005]
// The division r = num / den is the device's sequence: reciprocal seed, two Newton steps, quotient, one residual
// correction.  The hardware seed (MUFU.RCP64H, ~20 good bits) is not reproducible on the host: SEEDPERT scales a
// 24-bit seed so that runs with 1 - 2^-21, 1 and 1 + 2^-21 bracket whatever the hardware returns; the corrected
// quotient is the same for all of them except on near-ties of the division (last argument).
static const int NI = 16;
// build with -ffp-contract=off (the device code is compiled with -fmad=false)
static double C[17], ATC[17], TB[17];
static void init(void){ for(int k=0;k<=NI;k++){ double a = (M_PI/4)*k/NI; C[k]=tan(a); ATC[k]=a; }
  for(int k=0;k<NI;k++){ TB[k]=tan((M_PI/4)*(k+0.5)/NI);} }
static float TPERT = 1.0f;
static double SEEDPERT = 1.0;
static inline double dev_div(double num, double den){
  double y = (double)(float)(1.0/den) * SEEDPERT;            /* seed: ~2^-21 relative error at worst */
  y = fma(y, fma(-den, y, 1.0), y);
  y = fma(y, fma(-den, y, 1.0), y);
  double r = num * y;
  return fma(fma(-den, r, num), y, r);
}
static inline double fast_atan2(float yf, float xf){
  double x=xf,y=yf; double ax=fabs(x), ay=fabs(y);
  double hi = ax>ay?ax:ay, lo = ax>ay?ay:ax;
  double a;
  if (lo==0.0) a = 0.0; // includes (0,0)
  else {
    // interval: largest k with lo >= TB[k-1]*hi ... nearest breakpoint c_k = tan(k*pi/64)
    /* nearest breakpoint from a quadratic fit of atan, as common.cuh; the device computes t with __fdividef
       (2 ulp), so a neighbouring k may be picked next to a boundary: KSHIFT = -1/0/+1 forces that here */
    int k; { float lof=(float)lo, hif=(float)hi; float t = lof/hif; t = t*TPERT; k = (int)(t*(21.5615f + -5.5615f*t) + 0.5f); if(k>16)k=16; if(k<0)k=0; }
    double c=C[k];
    double r = dev_div(fma(-c,hi,lo), fma(c,lo,hi));
    double r2=r*r;
    double p = fma(r2,-1.0/11.0,1.0/9.0); p=fma(r2,p,-1.0/7.0); p=fma(r2,p,1.0/5.0); p=fma(r2,p,-1.0/3.0); p=p*r2;
    a = ATC[k] + fma(r,p,r);
  }
  if (ay>ax) a = M_PI_2 - a;
  if (signbit(x)) a = M_PI - a;
  return copysign(a, y);
}
int main(int argc,char**argv){ init(); if(argc>2) TPERT=(float)atof(argv[2]); if(argc>3) SEEDPERT=atof(argv[3]); long n = argc>1?atol(argv[1]):100000000; uint64_t s=88172645463325252ULL; long mism=0; double maxulp=0;
  for(long i=0;i<n;i++){ s^=s<<13; s^=s>>7; s^=s<<17; uint32_t a=(uint32_t)s, b=(uint32_t)(s>>32);
    float x,y; // mix of scales: gradients are differences of floats in [0,255]
    int mode = i&3;
    if(mode==0){ x=((int32_t)a)*(1.0f/8388608.0f); y=((int32_t)b)*(1.0f/8388608.0f);} // +-256 range
    else if(mode==1){ memcpy(&x,&a,4); memcpy(&y,&b,4); if(!isfinite(x)||!isfinite(y)||fabsf(x)>1e30f||fabsf(y)>1e30f||(fabsf(x)<1e-30f&&x!=0)||(fabsf(y)<1e-30f&&y!=0)) {x=1.5f;y=-2.5f;} }
    else if(mode==2){ x=((int32_t)a>>8)*(1.0f/65536.0f); y=((int32_t)(b>>20))*(1.0f/16.0f);} 
    else { x=(float)((int)(a%2001)-1000)*0.125f; y=(float)((int)(b%2001)-1000)*0.125f; }
    double ref=atan2((double)y,(double)x), got=fast_atan2(y,x);
    if((float)ref!=(float)got && !(isnan(ref)&&isnan(got))) { mism++; if(mism<5) printf("mismatch x=%a y=%a ref=%.17g got=%.17g\n",x,y,ref,got);} 
    if(ref!=0){ double u=fabs(got-ref)/ (fabs(ref)*2.220446049250313e-16); if(u>maxulp) maxulp=u; }
    else if (got!=0 || signbit(got)!=signbit(ref)) {mism++;}
  }
  printf("n=%ld float mismatches=%ld max err=%.2f double-ulps\n",n,mism,maxulp); return mism!=0; }
