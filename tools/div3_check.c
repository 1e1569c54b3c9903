// Exhaustive check (all 2^32 fp32 inputs) that x*(1/3) with one fma residual correction equals the IEEE double
// division x/3.0 after rounding to fp32 (used by k_orient's histogram smoothing, orientation_cpu.cl:108).
//   gcc -O2 -fopenmp -ffp-contract=off -mfma tools/div3_check.c -o tools/div3_check -lm && tools/div3_check
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
int main(void){ long long bad=0,badd=0;
#pragma omp parallel for reduction(+:bad,badd) schedule(static,1<<20)
 for(long long b=0;b<(1LL<<32);b++){ uint32_t u=(uint32_t)b; float x; memcpy(&x,&u,4); if(!(x==x)||isinf(x)) continue;
   double xd=(double)x; double ref=xd/3.0; const double C=1.0/3.0; double q0=xd*C; double r=fma(-3.0,q0,xd); double q1=fma(r,C,q0);
   if(memcmp(&ref,&q1,8)!=0){badd++; if((float)ref!=(float)q1) bad++;} }
 printf("double mismatches %lld float mismatches %lld\n",badd,bad); return bad!=0;}
