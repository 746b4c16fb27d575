// XLA-FFI handlers over the C ABI of liblarnd_b200.so (include/larnd_b200.h): the jax.ffi face of the drop-in boundary.
//
// The reference (pgranger23/larnd-sim-jax) has no FFI; its hot path is the JAX functions of src/larndsim/sim_jax.py.  Each
// handler below is the custom call that replaces the XLA lowering of one of them; ffi/sim_b200.py wraps them in
// jax.custom_vjp behind the reference's own signatures (simulate_wfs :689, simulate_stochastic :738, simulate_parametrized
// :339).  Build (done by __graft_entry__.build() as soon as `import jax` works):
//   g++ -std=c++17 -shared -fPIC -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") -Iinclude \
//       -I/usr/local/cuda/include ffi/larnd_ffi.cc -Llarnd-sim-jax_b200/larndsim_b200 -llarnd_b200 -lcudart -o ffi/liblarnd_ffi.so
// jax is absent from the build image and from the GPU box (profiles/r2_probe_jax.txt): there the file is only checked
// against ffi/mock/xla/ffi/api/ffi.h (tests/test_ffi_shim.py).
//
// Conventions shared by all handlers
//   * every operand / result is a DEVICE buffer owned by XLA; results are pre-allocated (static shapes), so the pixel
//     capacity Npix is chosen in Python (the reference's pad_size) and the true counts come back in `counts`;
//   * the parameter block larnd_params_t and the column map larnd_columns_t arrive as uint8 operands assembled INSIDE the
//     traced function (the fitted leaves are traced float32 scalars); the kernels take them by value, so the handler
//     copies the ~1 KB to the host and waits for it — the one host synchronisation per call (XLA's own lowering of
//     simulate_wfs synchronises at jnp.unique as well);
//   * the LUT handle (larnd_lut_create, made once per response_template by the Python side) travels as an int64 attribute;
//   * errors: the library's code + larnd_last_error() text as ffi::Error.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <string>

#include "larnd_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

ffi::Error Fail(int rc) {
  return ffi::Error(rc == LARND_E_ARG ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal,
                    std::string("larnd_b200: ") + larnd_last_error());
}

// device blobs -> host structs (see the conventions above)
ffi::Error FetchParams(cudaStream_t stream, const ffi::Buffer<ffi::U8>& pod, larnd_params_t* hp, const ffi::Buffer<ffi::U8>* cols,
                       larnd_columns_t* hc) {
  if (pod.element_count() != sizeof(larnd_params_t))
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "params blob size does not match larnd_params_t (ABI mismatch)");
  if (cols && cols->element_count() != sizeof(larnd_columns_t))
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "columns blob size does not match larnd_columns_t (ABI mismatch)");
  if (cudaMemcpyAsync(hp, pod.typed_data(), sizeof(*hp), cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
      (cols && cudaMemcpyAsync(hc, cols->typed_data(), sizeof(*hc), cudaMemcpyDeviceToHost, stream) != cudaSuccess) ||
      cudaStreamSynchronize(stream) != cudaSuccess)
    return ffi::Error(ffi::ErrorCode::kInternal, "copying the parameter block to the host failed");
  return ffi::Error::Success();
}

const larnd_lut_t* Lut(int64_t handle) { return reinterpret_cast<const larnd_lut_t*>(static_cast<intptr_t>(handle)); }

// ---- simulate_drift_new + jnp.unique (sim_jax.py:375-453,717): fills the workspace records, counts[0] = n_unique ------
ffi::Error LutPrepare(cudaStream_t stream, ffi::Buffer<ffi::F32> tracks, ffi::Buffer<ffi::U8> pod, ffi::Buffer<ffi::U8> cols,
                      int64_t lut, int32_t n_events, ffi::ResultBuffer<ffi::U8> ws, ffi::ResultBuffer<ffi::S32> counts) {
  larnd_params_t hp;
  larnd_columns_t hc;
  if (auto e = FetchParams(stream, pod, &hp, &cols, &hc); !e.success()) return e;
  const int64_t n = tracks.dimensions().size() ? tracks.dimensions()[0] : 0;
  if (cudaMemsetAsync(counts->typed_data(), 0, 4 * sizeof(int32_t), stream) != cudaSuccess)
    return ffi::Error(ffi::ErrorCode::kInternal, "cudaMemsetAsync failed");
  const int rc = larnd_lut_prepare(tracks.typed_data(), n, &hc, &hp, Lut(lut), n_events, ws->typed_data(), ws->size_bytes(),
                                   counts->typed_data(), stream);
  return rc ? Fail(rc) : ffi::Error::Success();
}

// ---- pad / sort / searchsorted + simulate_signals (sim_jax.py:718-736,142-286) on a prepared workspace ----------------
// ws and counts are aliased to the outputs of the same name (input_output_aliases on the Python side): the records stay in
// place for the backward call.  wfs is the FULL padded buffer (Npix, stride >= n_ticks): simulate_wfs returns [:, 1:n_ticks].
ffi::Error LutAccumulate(cudaStream_t stream, ffi::Buffer<ffi::U8> ws_in, ffi::Buffer<ffi::S32> counts_in, ffi::Buffer<ffi::U8> pod,
                         int64_t lut, int32_t n_events, int64_t n_segments, int32_t flags, ffi::ResultBuffer<ffi::F32> wfs,
                         ffi::ResultBuffer<ffi::S32> upix, ffi::ResultBuffer<ffi::U8> ws, ffi::ResultBuffer<ffi::S32> counts) {
  larnd_params_t hp;
  if (auto e = FetchParams(stream, pod, &hp, nullptr, nullptr); !e.success()) return e;
  if (ws->untyped_data() != ws_in.untyped_data() || counts->untyped_data() != counts_in.untyped_data())
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "LutAccumulate: ws / counts must be aliased to their outputs");
  const auto d = wfs->dimensions();
  if (d.size() != 2 || d[1] < hp.n_ticks) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "wfs must be (Npix, >= n_ticks)");
  const int rc = larnd_lut_accumulate(n_segments, &hp, Lut(lut), n_events, (int32_t)d[0], flags, ws->typed_data(), ws->size_bytes(),
                                      upix->typed_data(), wfs->typed_data(), d[1], counts->typed_data(), stream);
  return rc ? Fail(rc) : ffi::Error::Success();
}

// ---- jax.grad through simulate_wfs: g_wfs is the gradient of the FULL padded buffer (column 0 = garbage tick) ----------
ffi::Error LutBackward(cudaStream_t stream, ffi::Buffer<ffi::F32> g_wfs, ffi::Buffer<ffi::U8> ws_in, ffi::Buffer<ffi::S32> counts,
                       ffi::Buffer<ffi::U8> pod, int64_t lut, int32_t n_events, int64_t n_segments, int32_t flags,
                       ffi::ResultBuffer<ffi::F32> grad, ffi::ResultBuffer<ffi::U8> ws) {
  larnd_params_t hp;
  if (auto e = FetchParams(stream, pod, &hp, nullptr, nullptr); !e.success()) return e;
  if (ws->untyped_data() != ws_in.untyped_data())
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "LutBackward: ws must be aliased to its output (the call uses scratch inside it)");
  const auto d = g_wfs.dimensions();
  if (d.size() != 2 || grad->element_count() != LARND_NPARAMS) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "LutBackward: bad shapes");
  if (cudaMemsetAsync(grad->typed_data(), 0, LARND_NPARAMS * sizeof(float), stream) != cudaSuccess)
    return ffi::Error(ffi::ErrorCode::kInternal, "cudaMemsetAsync failed");
  const int rc = larnd_lut_backward(n_segments, &hp, Lut(lut), n_events, (int32_t)d[0], flags, ws->typed_data(), ws->size_bytes(),
                                    counts.typed_data(), g_wfs.typed_data(), d[1], grad->typed_data(), stream);
  return rc ? Fail(rc) : ffi::Error::Success();
}

// ---- get_adc_values + digitize + id2pixel + get_pixel_coordinates + get_hit_z, dense (Npix, 10) outputs -----------------
// wfs is the FULL padded buffer again; the kernel reads columns 1 .. n_ticks-1.  noise: 0 elements (noise-free) or the
// Npix * 31 standard normals of larnd_rng_fee_noise's layout.  parse_output (sim_jax.py:620-647) stays jnp code on the dense
// outputs, like in the reference.
ffi::Error FeeForward(cudaStream_t stream, ffi::Buffer<ffi::F32> wfs, ffi::Buffer<ffi::S32> upix, ffi::Buffer<ffi::U8> pod,
                      ffi::Buffer<ffi::F32> noise, ffi::ResultBuffer<ffi::F32> adc, ffi::ResultBuffer<ffi::F32> ticks,
                      ffi::ResultBuffer<ffi::F32> pixel_z, ffi::ResultBuffer<ffi::F32> pixel_x, ffi::ResultBuffer<ffi::F32> pixel_y,
                      ffi::ResultBuffer<ffi::S32> event, ffi::ResultBuffer<ffi::F32> saved, ffi::ResultBuffer<ffi::S32> n_valid,
                      ffi::ResultBuffer<ffi::U8> scratch) {
  larnd_params_t hp;
  if (auto e = FetchParams(stream, pod, &hp, nullptr, nullptr); !e.success()) return e;
  const auto d = wfs.dimensions();
  if (d.size() != 2 || d[1] < hp.n_ticks) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "wfs must be (Npix, >= n_ticks)");
  const int32_t npix = (int32_t)d[0];
  if (scratch->size_bytes() < larnd_fee_scratch_bytes(npix)) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "FeeForward: scratch too small");
  const float* nz = noise.element_count() ? noise.typed_data() : nullptr;
  const int rc = larnd_fee_forward(wfs.typed_data() + 1, d[1], upix.typed_data(), npix, &hp, nz, adc->typed_data(), ticks->typed_data(),
                                   pixel_z->typed_data(), pixel_x->typed_data(), pixel_y->typed_data(), event->typed_data(),
                                   saved->typed_data(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                   n_valid->typed_data(), scratch->typed_data(), scratch->size_bytes(), stream);
  return rc ? Fail(rc) : ffi::Error::Success();
}

// ---- jax.grad through get_adc_values + digitize: g_adc (Npix, 10) -> gradient of the FULL padded waveform buffer -------
ffi::Error FeeBackward(cudaStream_t stream, ffi::Buffer<ffi::F32> g_adc, ffi::Buffer<ffi::F32> ticks, ffi::Buffer<ffi::F32> saved,
                       ffi::Buffer<ffi::U8> pod, int32_t raw_charge, ffi::ResultBuffer<ffi::F32> g_wfs) {
  larnd_params_t hp;
  if (auto e = FetchParams(stream, pod, &hp, nullptr, nullptr); !e.success()) return e;
  const auto d = g_wfs->dimensions();
  if (d.size() != 2 || d[1] < hp.n_ticks) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "g_wfs must be (Npix, >= n_ticks)");
  if (cudaMemsetAsync(g_wfs->typed_data(), 0, g_wfs->size_bytes(), stream) != cudaSuccess)   // column 0 and the padding
    return ffi::Error(ffi::ErrorCode::kInternal, "cudaMemsetAsync failed");
  const int rc = larnd_fee_backward(g_adc.typed_data(), ticks.typed_data(), saved.typed_data(), (int32_t)d[0], &hp,
                                    g_wfs->typed_data() + 1, d[1], raw_charge, stream);
  return rc ? Fail(rc) : ffi::Error::Success();
}

// ---- MC-current mode: simulate_drift(mc_diff) + current_mc + accumulate_signals_parametrized (sim_jax.py:339-372) --------
ffi::Error McForward(cudaStream_t stream, ffi::Buffer<ffi::F32> tracks, ffi::Buffer<ffi::F32> rnd, ffi::Buffer<ffi::U8> pod,
                     ffi::Buffer<ffi::U8> cols, int32_t n_events, ffi::ResultBuffer<ffi::F32> wfs, ffi::ResultBuffer<ffi::S32> upix,
                     ffi::ResultBuffer<ffi::S32> counts, ffi::ResultBuffer<ffi::U8> ws) {
  larnd_params_t hp;
  larnd_columns_t hc;
  if (auto e = FetchParams(stream, pod, &hp, &cols, &hc); !e.success()) return e;
  const auto d = wfs->dimensions();
  if (d.size() != 2 || d[1] != hp.n_ticks) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "MC wfs must be (Npix, n_ticks)");
  if (cudaMemsetAsync(counts->typed_data(), 0, 4 * sizeof(int32_t), stream) != cudaSuccess)
    return ffi::Error(ffi::ErrorCode::kInternal, "cudaMemsetAsync failed");
  const int64_t n = tracks.dimensions()[0];
  const int rc = larnd_mc_forward(tracks.typed_data(), n, &hc, &hp, rnd.typed_data(), n_events, (int32_t)d[0], ws->typed_data(),
                                  ws->size_bytes(), upix->typed_data(), wfs->typed_data(), counts->typed_data(), stream);
  return rc ? Fail(rc) : ffi::Error::Success();
}

ffi::Error McBackward(cudaStream_t stream, ffi::Buffer<ffi::F32> g_wfs, ffi::Buffer<ffi::F32> tracks, ffi::Buffer<ffi::F32> rnd,
                      ffi::Buffer<ffi::U8> ws_in, ffi::Buffer<ffi::S32> counts, ffi::Buffer<ffi::U8> pod, ffi::Buffer<ffi::U8> cols,
                      int32_t n_events, ffi::ResultBuffer<ffi::F32> grad, ffi::ResultBuffer<ffi::U8> ws) {
  larnd_params_t hp;
  larnd_columns_t hc;
  if (auto e = FetchParams(stream, pod, &hp, &cols, &hc); !e.success()) return e;
  if (ws->untyped_data() != ws_in.untyped_data()) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "McBackward: ws must be aliased");
  const auto d = g_wfs.dimensions();
  if (d.size() != 2 || grad->element_count() != LARND_NPARAMS) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "McBackward: bad shapes");
  if (cudaMemsetAsync(grad->typed_data(), 0, LARND_NPARAMS * sizeof(float), stream) != cudaSuccess)
    return ffi::Error(ffi::ErrorCode::kInternal, "cudaMemsetAsync failed");
  const int rc = larnd_mc_backward(tracks.typed_data(), tracks.dimensions()[0], &hc, &hp, rnd.typed_data(), n_events, (int32_t)d[0],
                                   ws->typed_data(), ws->size_bytes(), counts.typed_data(), g_wfs.typed_data(), d[1],
                                   grad->typed_data(), stream);
  return rc ? Fail(rc) : ffi::Error::Success();
}

}  // namespace

using F32 = ffi::Buffer<ffi::F32>;
using S32 = ffi::Buffer<ffi::S32>;
using U8 = ffi::Buffer<ffi::U8>;

XLA_FFI_DEFINE_HANDLER_SYMBOL(larnd_ffi_lut_prepare, LutPrepare,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F32>().Arg<U8>().Arg<U8>()
                                  .Attr<int64_t>("lut").Attr<int32_t>("n_events").Ret<U8>().Ret<S32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(larnd_ffi_lut_accumulate, LutAccumulate,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<U8>().Arg<S32>().Arg<U8>()
                                  .Attr<int64_t>("lut").Attr<int32_t>("n_events").Attr<int64_t>("n_segments").Attr<int32_t>("flags")
                                  .Ret<F32>().Ret<S32>().Ret<U8>().Ret<S32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(larnd_ffi_lut_backward, LutBackward,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F32>().Arg<U8>().Arg<S32>().Arg<U8>()
                                  .Attr<int64_t>("lut").Attr<int32_t>("n_events").Attr<int64_t>("n_segments").Attr<int32_t>("flags")
                                  .Ret<F32>().Ret<U8>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(larnd_ffi_fee_forward, FeeForward,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F32>().Arg<S32>().Arg<U8>().Arg<F32>()
                                  .Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<S32>().Ret<F32>().Ret<S32>().Ret<U8>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(larnd_ffi_fee_backward, FeeBackward,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<U8>()
                                  .Attr<int32_t>("raw_charge").Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(larnd_ffi_mc_forward, McForward,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F32>().Arg<F32>().Arg<U8>().Arg<U8>()
                                  .Attr<int32_t>("n_events").Ret<F32>().Ret<S32>().Ret<S32>().Ret<U8>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(larnd_ffi_mc_backward, McBackward,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<U8>().Arg<S32>()
                                  .Arg<U8>().Arg<U8>().Attr<int32_t>("n_events").Ret<F32>().Ret<U8>());
