// oracle/tr1_order.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// The reference's v3 transform keeps its result in a
//   std::tr1::unordered_map<int, complex_t>          (src/sfft.h:35-36)
// and twice walks that map to peel known coefficients out of freshly filled
// buckets (src/computefourier-3.0.cc:881-887, :940-947).  The WALK ORDER decides
// the order of floating-point subtractions, so a bit-exact restatement needs the
// very same container.  This file wraps it behind a C interface for sfft_oracle.c.
#include <tr1/unordered_map>

struct orc_val { double re, im; };
typedef std::tr1::unordered_map<int, orc_val, std::tr1::hash<int> > map_t;

extern "C" {

void *orc_map_new(void) { return new map_t(); }
void orc_map_free(void *m) { delete (map_t *)m; }
// operator[]: inserts a zero entry when the key is absent
double *orc_map_ref(void *m, int key)
{
  map_t &mm = *(map_t *)m;
  map_t::iterator it = mm.find(key);
  if (it == mm.end()) {
    orc_val z; z.re = 0; z.im = 0;
    it = mm.insert(std::make_pair(key, z)).first;
  }
  return &it->second.re;
}
int orc_map_size(void *m) { return (int)((map_t *)m)->size(); }
// keys/values in iteration order
void orc_map_dump(void *m, int *keys, double *vals)
{
  map_t &mm = *(map_t *)m;
  int i = 0;
  for (map_t::iterator it = mm.begin(); it != mm.end(); ++it, ++i) {
    keys[i] = it->first;
    vals[2 * i] = it->second.re;
    vals[2 * i + 1] = it->second.im;
  }
}

}
