#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
static double LT[128][2];  // 1/c_i (rounded), log(c_i)
static double ET[64];      // 2^(k/64)
static inline uint64_t d2u(double x){uint64_t u;memcpy(&u,&x,8);return u;}
static inline double u2d(uint64_t u){double x;memcpy(&x,&u,8);return x;}
static void init(){
  for(int i=0;i<128;i++){ long double c=1.0L+(i+0.5L)/128.0L; double rc=(double)(1.0L/c);
    /* log of the value actually divided out: 1/rc, so that m*rc - 1 is consistent */
    LT[i][0]=rc; LT[i][1]=(double)(-logl((long double)rc)); }
  for(int k=0;k<64;k++) ET[k]=(double)powl(2.0L,k/64.0L);
}
static inline double fast_log(double x){
  uint64_t u=d2u(x); int hi=(int)(u>>32);
  int e=(hi>>20)-1023; int i=(hi>>13)&127;
  double m=u2d((u&0x000fffffffffffffull)|0x3ff0000000000000ull);
  double r=fma(m,LT[i][0],-1.0);
  double p=fma(r,-1.0/6,1.0/5); p=fma(r,p,-0.25); p=fma(r,p,1.0/3); p=fma(r,p,-0.5);
  double r2=r*r; double l1=fma(r2,p,r);
  const double LN2H=0x1.62e42fefa38p-1, LN2L=0x1.ef35793c7673p-45;
  double ed=(double)e;
  double s=fma(ed,LN2H,LT[i][1]);   // exact-ish: LN2H has 42 bits
  return (s+l1)+ed*LN2L;
}
static inline double fast_exp(double z){
  const double INV=64.0/0.6931471805599453094172321; 
  const double LH=0x1.62e42fefa38p-7, LL=0x1.ef35793c7673p-51; // ln2/64 hi, lo
  double kd=rint(z*INV); 
  double r=fma(-kd,LH,z); r=fma(-kd,LL,r);
  long k=(long)kd;
  double p=fma(r,1.0/120,1.0/24); p=fma(r,p,1.0/6); p=fma(r,p,0.5);
  double r2=r*r; double q=fma(r2,p,r);
  double T=ET[k&63]; 
  double v=fma(T,q,T);
  uint64_t u=d2u(v); u+= (uint64_t)((k>>6))<<52; return u2d(u);
}
static inline double fast_pw(double x,double y){ return fast_exp(y*fast_log(x)); }
int main(){
  init(); srand(1);
  double worst=0, worst_std=0; double wx=0,wy=0;
  for(long n=0;n<20000000;n++){
    double lx=-8+ (rand()/(double)RAND_MAX)*(20); // x in 1e-8..1e12
    double x=pow(10.0,lx); double y=-10+(rand()/(double)RAND_MAX)*62; 
    long double ref=powl((long double)x,(long double)y);
    if(!(ref>1e-290L && ref<1e290L)) continue;
    double a=fast_pw(x,y), b=exp(y*log(x));
    double ea=fabs((double)((a-ref)/ref)), eb=fabs((double)((b-ref)/ref));
    if(ea>worst){worst=ea;wx=x;wy=y;} if(eb>worst_std)worst_std=eb;
  }
  printf("fast_pw worst rel %.3e at x=%g y=%g ; exp(y log x) worst %.3e\n",worst,wx,wy,worst_std);
  /* log / exp alone */
  double wl=0,we=0;
  for(long n=0;n<20000000;n++){ double x=pow(10.0,-300+600*(rand()/(double)RAND_MAX)); long double r=logl(x); double e=fabs((double)((fast_log(x)-r)/r)); if(fabs((double)r)>1e-3 && e>wl)wl=e;
    double z=-700+1400*(rand()/(double)RAND_MAX); long double re=expl(z); double ee=fabs((double)((fast_exp(z)-re)/re)); if(ee>we)we=ee; }
  printf("log worst rel %.3e exp worst rel %.3e\n",wl,we);
  return 0;
}
