// Instantiations of sweep_prod_kernel<float, NFULL, TAIL, NK_RULE_EXCHANGE, MULTI=true> (one translation unit per variant: parallel build).
#include "sweep_prod.cuh"

namespace nk {

int launch_prod_f32_exchange_multi(cudaStream_t stream, const ProdArgs &a, int nfull, int tail) {
#define NK_PROD_CASE(NF, TL) \
  if (nfull == NF && tail == TL) return launch_prod<float, NF, TL, NK_RULE_EXCHANGE, true>(stream, a);
  NK_PROD_CASE(2, 0)
  NK_PROD_CASE(2, 1)
  NK_PROD_CASE(2, 2)
  NK_PROD_CASE(3, 0)
  NK_PROD_CASE(3, 1)
  NK_PROD_CASE(3, 2)
  NK_PROD_CASE(4, 0)
  NK_PROD_CASE(5, 0)
#undef NK_PROD_CASE
  set_error("sweep_prod: no instantiation for this number of hidden units");
  return NK_EUNSUPPORTED;
}

}  // namespace nk
