// Host build of the Rayleigh-Ritz core (lsa_fw_b200/csrc/rr_core.h) with a team of one thread.
// TEST INFRASTRUCTURE: lets the CPU suite check the dense Schur / ordering / restart logic that the
// CUDA kernel k_rr executes, without a GPU.  Not linked into the product library.
#include <vector>

#include "../../lsa_fw_b200/csrc/rr_core.h"

extern "C" int rr_host_full(int m, int ld, int nconv, int nev, int which, int transform, int last, double tol,
                            double sigma_re, double sigma_im, double beta_scale, double* S, double* Q,
                            double* theta, double* resid, int* out4) {
  using namespace lsa;
  RrParams p;
  p.m = m; p.ld = ld; p.ldq = m; p.nconv = nconv; p.nev = nev; p.which = which; p.transform = transform;
  p.last = last; p.tol = tol; p.sigma = mk(sigma_re, sigma_im); p.beta_scale = beta_scale;
  std::vector<double> rot_c(m + 1), key(m + 1);
  std::vector<z128> rot_s(m + 1), vec(m + 1), brow(m + 1), ywork((size_t)(nev + 8) * m + 1);
  int iflag[4] = {0, 0, 0, 0};
  RrWork w{rot_c.data(), rot_s.data(), vec.data(), key.data(), iflag};
  RrOut out{};
  rr_full((z128*)S, (z128*)Q, p, (z128*)theta, resid, brow.data(), ywork.data(), &out, 0, 1, w);
  out4[0] = out.nconv; out4[1] = out.keep; out4[2] = out.status; out4[3] = 0;
  return 0;
}

extern "C" int rr_host_schur(int m, int ld, int lo, double* S, double* Q) {
  using namespace lsa;
  std::vector<double> rot_c(m + 1), key(m + 1);
  std::vector<z128> rot_s(m + 1), vec(m + 1);
  int iflag[4] = {0, 0, 0, 0};
  RrWork w{rot_c.data(), rot_s.data(), vec.data(), key.data(), iflag};
  return rr_schur((z128*)S, ld, (z128*)Q, m, m, lo, 0, 1, w);
}
