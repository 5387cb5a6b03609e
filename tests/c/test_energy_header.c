#include <stdio.h>
#include <string.h>
#include "wn_energy.h"
double wno_compute_energy(double r, double cosine);
int main(void){
  int bad=0,n=0; wn_energy_params p=wn_energy_defaults();
  for (float r=2.0f; r<7.2f; r+=0.0371f) for (float c=-1.0f; c<=1.0f; c+=0.0173f){
    float a=wn_edge_energy(r,c,&p); float b=(float)wno_compute_energy((double)r,(double)c);
    if (memcmp(&a,&b,4)) ++bad; ++n;
  }
  printf("%d of %d differ\n",bad,n); return bad!=0;
}
