// jax.ffi registration shim for libnkb200 (NOT built in this repository: the XLA FFI headers ship with jaxlib,
// which is not installed in this image; build.py compiles only netket_b200/csrc/*.cu).
//
// Build where jax is available:
//   g++ -O2 -fPIC -shared -std=c++17 -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") \
//       -I include -I /usr/local/cuda/include nkb200_jax_ffi.cc -L netket_b200/lib -lnkb200 -o libnkb200_jax.so
// and register from Python:
//   jax.ffi.register_ffi_target("nkb200_sweep", jax.ffi.pycapsule(lib.NkSweep), platform="CUDA")
//
// XLA-FFI rules honoured: buffers are XLA-owned device pointers, outputs pre-allocated by XLA, the work
// is enqueued on the stream XLA provides and never synchronises, errors come back as ffi::Error, no mutable globals.
#include <cuda_runtime.h>

#include "nkb200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static int dtype_code(ffi::DataType t) { return t == ffi::DataType::F32 ? NK_F32 : NK_F64; }

static ffi::Error fail() { return ffi::Error(ffi::ErrorCode::kInternal, nk_last_error()); }

// samples, log_prob, sigma', n_accepted', E_loc, statistics sums  =  sweep(W, b, a, sigma, n_accepted, edges; attrs)
static ffi::Error SweepImpl(cudaStream_t stream, ffi::AnyBuffer W, ffi::AnyBuffer b, ffi::AnyBuffer a, ffi::Buffer<ffi::S8> sigma,
                            ffi::Buffer<ffi::S64> n_accepted, ffi::Buffer<ffi::S32> edges, ffi::Result<ffi::Buffer<ffi::S8>> samples,
                            ffi::Result<ffi::AnyBuffer> log_prob, ffi::Result<ffi::Buffer<ffi::S8>> sigma_out,
                            ffi::Result<ffi::Buffer<ffi::S64>> n_accepted_out, ffi::Result<ffi::AnyBuffer> eloc,
                            ffi::Result<ffi::Buffer<ffi::F64>> sums, ffi::Result<ffi::Buffer<ffi::U8>> workspace, int32_t rule,
                            int32_t chain_length, int32_t n_discard, int32_t sweep_size, double machine_pow, double h, double J,
                            double stats_shift, uint64_t seed, uint64_t t, uint64_t chain_offset) {
  const auto dims = sigma.dimensions();
  const int64_t B = dims[0];
  const int32_t N = (int32_t)dims[1];
  nk_rbm_t rbm{W.untyped_data(), b.untyped_data(), a.untyped_data(), N, (int32_t)W.dimensions()[1], dtype_code(W.element_type()), 0};
  // functional semantics: XLA gives fresh output buffers; copy the state into them and update in place
  cudaMemcpyAsync(sigma_out->typed_data(), sigma.typed_data(), (size_t)B * N, cudaMemcpyDeviceToDevice, stream);
  cudaMemcpyAsync(n_accepted_out->typed_data(), n_accepted.typed_data(), (size_t)B * 8, cudaMemcpyDeviceToDevice, stream);
  nk_chains_t ch{sigma_out->typed_data(), log_prob->untyped_data(), n_accepted_out->typed_data(), workspace->typed_data(), B, seed, t,
                 chain_offset};
  nk_ising_t ising{edges.typed_data(), (int32_t)edges.dimensions()[0], 0, h, J};
  nk_sweep_t args{};
  args.rule = rule;
  args.chain_length = chain_length;
  args.n_discard = n_discard;
  args.sweep_size = sweep_size;
  args.machine_pow = machine_pow;
  args.samples_out = samples->typed_data();
  args.ising = &ising;
  args.eloc_out = eloc->untyped_data();
  args.eloc_dtype = dtype_code(eloc->element_type());
  args.path = NK_PATH_AUTO;
  args.stats_out = sums->typed_data();  // NK_STATS_NPARTIAL doubles: psum over the mesh, then nk_stats_finalize's arithmetic
  args.stats_shift = stats_shift;
  return nk_sweep(stream, &rbm, &ch, &args) == NK_OK ? ffi::Error::Success() : fail();
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(NkSweep, SweepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()          // W
                                  .Arg<ffi::AnyBuffer>()          // b
                                  .Arg<ffi::AnyBuffer>()          // a
                                  .Arg<ffi::Buffer<ffi::S8>>()    // sigma
                                  .Arg<ffi::Buffer<ffi::S64>>()   // n_accepted
                                  .Arg<ffi::Buffer<ffi::S32>>()   // edges
                                  .Ret<ffi::Buffer<ffi::S8>>()    // samples (B, chain_length, N)
                                  .Ret<ffi::AnyBuffer>()          // log_prob (B,)
                                  .Ret<ffi::Buffer<ffi::S8>>()    // sigma'
                                  .Ret<ffi::Buffer<ffi::S64>>()   // n_accepted'
                                  .Ret<ffi::AnyBuffer>()          // E_loc (B, chain_length)
                                  .Ret<ffi::Buffer<ffi::F64>>()   // shifted statistics sums (NK_STATS_NPARTIAL)
                                  .Ret<ffi::Buffer<ffi::U8>>()    // workspace (nk_sweep_workspace_bytes)
                                  .Attr<int32_t>("rule")
                                  .Attr<int32_t>("chain_length")
                                  .Attr<int32_t>("n_discard")
                                  .Attr<int32_t>("sweep_size")
                                  .Attr<double>("machine_pow")
                                  .Attr<double>("h")
                                  .Attr<double>("J")
                                  .Attr<double>("stats_shift")
                                  .Attr<uint64_t>("seed")
                                  .Attr<uint64_t>("t")
                                  .Attr<uint64_t>("chain_offset"));

// xp, mels = IsingJax.get_conn_padded(x)
static ffi::Error IsingConnImpl(cudaStream_t stream, ffi::Buffer<ffi::S8> x, ffi::Buffer<ffi::S32> edges,
                                ffi::Result<ffi::Buffer<ffi::S8>> xp, ffi::Result<ffi::AnyBuffer> mels, double h, double J) {
  const auto d = x.dimensions();
  nk_ising_t op{edges.typed_data(), (int32_t)edges.dimensions()[0], 0, h, J};
  const int rc = nk_ising_conn(stream, &op, x.typed_data(), d[0], (int32_t)d[1], xp->typed_data(), mels->untyped_data(),
                               dtype_code(mels->element_type()));
  return rc == NK_OK ? ffi::Error::Success() : fail();
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(NkIsingConn, IsingConnImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::S8>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::S8>>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Attr<double>("h")
                                  .Attr<double>("J"));

// logpsi = RBM.apply(params, sigma)
static ffi::Error LogPsiImpl(cudaStream_t stream, ffi::AnyBuffer W, ffi::AnyBuffer b, ffi::AnyBuffer a, ffi::Buffer<ffi::S8> sigma,
                             ffi::Result<ffi::AnyBuffer> out) {
  const auto d = sigma.dimensions();
  nk_rbm_t rbm{W.untyped_data(), b.untyped_data(), a.untyped_data(), (int32_t)d[1], (int32_t)W.dimensions()[1],
               dtype_code(W.element_type()), 0};
  return nk_rbm_logpsi(stream, &rbm, sigma.typed_data(), d[0], out->untyped_data(), nullptr) == NK_OK ? ffi::Error::Success() : fail();
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(NkRbmLogPsi, LogPsiImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::Buffer<ffi::S8>>()
                                  .Ret<ffi::AnyBuffer>());

// E_loc = local_value_kernel for (RBM, Ising) on given configurations (stand-alone local estimator, seam S4)
static ffi::Error ElocIsingImpl(cudaStream_t stream, ffi::AnyBuffer W, ffi::AnyBuffer b, ffi::AnyBuffer a, ffi::Buffer<ffi::S8> sigma,
                                ffi::Buffer<ffi::S32> edges, ffi::Result<ffi::AnyBuffer> eloc, ffi::Result<ffi::Buffer<ffi::U8>> workspace,
                                double h, double J) {
  const auto d = sigma.dimensions();
  nk_rbm_t rbm{W.untyped_data(), b.untyped_data(), a.untyped_data(), (int32_t)d[1], (int32_t)W.dimensions()[1],
               dtype_code(W.element_type()), 0};
  nk_ising_t op{edges.typed_data(), (int32_t)edges.dimensions()[0], 0, h, J};
  const int rc = nk_eloc_ising_rbm(stream, &rbm, &op, sigma.typed_data(), d[0], eloc->untyped_data(), dtype_code(eloc->element_type()),
                                   NK_PATH_AUTO, workspace->typed_data());
  return rc == NK_OK ? ffi::Error::Success() : fail();
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(NkElocIsing, ElocIsingImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()          // W
                                  .Arg<ffi::AnyBuffer>()          // b
                                  .Arg<ffi::AnyBuffer>()          // a
                                  .Arg<ffi::Buffer<ffi::S8>>()    // sigma (B, N)
                                  .Arg<ffi::Buffer<ffi::S32>>()   // edges (E, 2)
                                  .Ret<ffi::AnyBuffer>()          // E_loc (B,)
                                  .Ret<ffi::Buffer<ffi::U8>>()    // workspace (nk_sweep_workspace_bytes)
                                  .Attr<double>("h")
                                  .Attr<double>("J"));

// sums[N*M + M + N] = sum_s dlogpsi(sigma_s) * (eloc_s - mean)   (forces before the cross-device psum and the 1/n scale)
static ffi::Error ForcesImpl(cudaStream_t stream, ffi::AnyBuffer W, ffi::AnyBuffer b, ffi::AnyBuffer a, ffi::Buffer<ffi::S8> samples,
                             ffi::AnyBuffer eloc, ffi::Result<ffi::Buffer<ffi::F64>> sums, ffi::Result<ffi::Buffer<ffi::U8>> workspace,
                             double mean) {
  const auto d = samples.dimensions();
  nk_rbm_t rbm{W.untyped_data(), b.untyped_data(), a.untyped_data(), (int32_t)d[1], (int32_t)W.dimensions()[1],
               dtype_code(W.element_type()), 0};
  const int rc = nk_forces_rbm(stream, &rbm, samples.typed_data(), d[0], eloc.untyped_data(), dtype_code(eloc.element_type()), mean,
                               sums->typed_data(), workspace->typed_data(), /*tanh_theta=*/nullptr);
  return rc == NK_OK ? ffi::Error::Success() : fail();
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(NkForces, ForcesImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()          // W
                                  .Arg<ffi::AnyBuffer>()          // b
                                  .Arg<ffi::AnyBuffer>()          // a
                                  .Arg<ffi::Buffer<ffi::S8>>()    // samples (n_s, N)
                                  .Arg<ffi::AnyBuffer>()          // E_loc (n_s,)
                                  .Ret<ffi::Buffer<ffi::F64>>()   // sums
                                  .Ret<ffi::Buffer<ffi::U8>>()    // workspace (nk_forces_workspace_bytes)
                                  .Attr<double>("mean"));

// OnlineStats.update: the eight state arrays of the pytree in, the eight updated arrays out (`_update_arrays`,
// netket/_src/stats/online_stats/kernels.py:116-190).  max_lag comes from the shapes, buf_len / decay are attributes.
static ffi::Error OnlineUpdateImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> count, ffi::Buffer<ffi::F64> mean, ffi::Buffer<ffi::F64> M2,
                                   ffi::Buffer<ffi::F64> cross, ffi::Buffer<ffi::F64> m1, ffi::Buffer<ffi::F64> m2,
                                   ffi::Buffer<ffi::F64> pairs, ffi::Buffer<ffi::F64> buf, ffi::AnyBuffer data,
                                   ffi::Result<ffi::Buffer<ffi::F64>> count_o, ffi::Result<ffi::Buffer<ffi::F64>> mean_o,
                                   ffi::Result<ffi::Buffer<ffi::F64>> M2_o, ffi::Result<ffi::Buffer<ffi::F64>> cross_o,
                                   ffi::Result<ffi::Buffer<ffi::F64>> m1_o, ffi::Result<ffi::Buffer<ffi::F64>> m2_o,
                                   ffi::Result<ffi::Buffer<ffi::F64>> pairs_o, ffi::Result<ffi::Buffer<ffi::F64>> buf_o, int32_t buf_len,
                                   double decay) {
  const int64_t n_chains = data.dimensions()[0], n = data.dimensions()[1];
  const int32_t max_lag = (int32_t)buf.dimensions()[1];
  nk_online_stats_t in{count.typed_data(), mean.typed_data(), M2.typed_data(), cross.typed_data(), m1.typed_data(),
                       m2.typed_data(),    pairs.typed_data(), buf.typed_data(), n_chains,          max_lag,
                       buf_len};
  nk_online_stats_t out{count_o->typed_data(), mean_o->typed_data(), M2_o->typed_data(), cross_o->typed_data(), m1_o->typed_data(),
                        m2_o->typed_data(),    pairs_o->typed_data(), buf_o->typed_data(), n_chains,            max_lag,
                        buf_len};
  const int rc = nk_online_stats_update(stream, &in, &out, data.untyped_data(), dtype_code(data.element_type()), n, decay);
  return rc == NK_OK ? ffi::Error::Success() : fail();
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(NkOnlineStatsUpdate, OnlineUpdateImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()   // _chain_count
                                  .Arg<ffi::Buffer<ffi::F64>>()   // _chain_mean
                                  .Arg<ffi::Buffer<ffi::F64>>()   // _chain_M2
                                  .Arg<ffi::Buffer<ffi::F64>>()   // _cross_sum
                                  .Arg<ffi::Buffer<ffi::F64>>()   // _m1_sum
                                  .Arg<ffi::Buffer<ffi::F64>>()   // _m2_sum
                                  .Arg<ffi::Buffer<ffi::F64>>()   // _pair_count
                                  .Arg<ffi::Buffer<ffi::F64>>()   // _chain_buf
                                  .Arg<ffi::AnyBuffer>()          // data (n_chains, n)
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Attr<int32_t>("buf_len")
                                  .Attr<double>("decay"));
