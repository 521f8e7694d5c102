// mrefsr_b200/csrc/torch_ext/deform_conv_ext.cpp -- the torch extension module that REPLACES the reference's
// `basicsr.ops.dcn.deform_conv_ext` build (setup.py:121-125; JIT-loaded by basicsr/ops/dcn/deform_conv.py:10-21).
//
// Same module name, same five pybind exports, same argument lists and by-value at::Tensor parameters as
// basicsr/ops/dcn/src/deform_conv_ext.cpp:150-164, so `from . import deform_conv_ext` in the reference's
// deform_conv.py binds to this file unchanged.  It is a thin shim: argument checks with the reference's messages
// (deform_conv_cuda.cpp:497-516, deform_conv_ext.cpp:124,146), a DeviceGuard, the caller's current stream, a workspace
// from torch's caching allocator -- and one call into the C ABI of libmrefsr_b200.so (include/mrefsr_b200.h), where the
// sm_100a kernels live.  No torch types cross that boundary.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "../../../include/mrefsr_b200.h"

namespace {

struct Workspace {
  at::Tensor buf;
  void* ptr;
  size_t bytes;
};

Workspace make_workspace(size_t bytes, const at::Tensor& like) {
  Workspace w;
  w.buf = at::empty({static_cast<int64_t>(bytes + 1024)}, like.options().dtype(at::kByte));
  auto base = reinterpret_cast<uintptr_t>(w.buf.data_ptr());
  auto aligned = (base + 1023) / 1024 * 1024;
  w.ptr = reinterpret_cast<void*>(aligned);
  w.bytes = bytes + 1024 - (aligned - base);
  return w;
}

// Error messages are formatted here and handed to TORCH_CHECK as ONE C string: with this image's torch 2.11 headers and
// gcc 13 the variadic c10::str(...) path of TORCH_CHECK crashes inside an extension module (reproduced with a
// three-line module), the single-string path does not.
[[noreturn]] void fail(const char* fmt, ...) {
  static thread_local char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  TORCH_CHECK(false, buf);
  abort();  // not reached
}

void check_rc(int rc, const char* what) {
  if (rc != 0) fail("%s failed (%d): %s", what, rc, mrefsr_last_error());
}

void check_kernel(int kernel_h, int kernel_w, int kh, int kw, int C, int channels_kernel, int group) {
  if (kh != kernel_h || kw != kernel_w)   // deform_conv_cuda.cpp:511-513
    fail("Input shape and kernel shape won't match: (%d x %d vs %d x %d).", kernel_h, kernel_w, kh, kw);
  if (C != channels_kernel * group)       // deform_conv_cuda.cpp:514-516
    fail("Input shape and kernel channels won't match: (%d vs %d).", C, channels_kernel * group);
}

// fp32 contiguous view of a tensor the kernels read (the reference dispatches on the scalar type; the sm_100a kernels
// compute in fp32 and other types are converted at this boundary)
at::Tensor f32(const at::Tensor& t) { return t.scalar_type() == at::kFloat ? t.contiguous() : t.to(at::kFloat).contiguous(); }

void write_back(at::Tensor dst, const at::Tensor& src) {
  if (dst.data_ptr() != src.data_ptr()) dst.view(src.sizes()).copy_(src);
}

// The exported functions have internal linkage: the reference's own extension defines functions with exactly these
// names and signatures, and two such modules in one process (the parity tests load both) must not interpose each other.
void modulated_deform_conv_forward(at::Tensor input, at::Tensor weight, at::Tensor bias, at::Tensor ones, at::Tensor offset,
                                   at::Tensor mask, at::Tensor output, at::Tensor columns, int kernel_h, int kernel_w,
                                   const int stride_h, const int stride_w, const int pad_h, const int pad_w,
                                   const int dilation_h, const int dilation_w, const int group, const int deformable_group,
                                   const bool with_bias) {
  TORCH_CHECK(input.is_cuda(), "modulated deform conv is not implemented on CPU");
  TORCH_CHECK(input.is_contiguous(), "input tensor has to be contiguous");
  TORCH_CHECK(weight.is_contiguous(), "weight tensor has to be contiguous");
  at::DeviceGuard guard(input.device());
  const int B = input.size(0), C = input.size(1), H = input.size(2), W = input.size(3);
  const int Co = weight.size(0), channels_kernel = weight.size(1), kh = weight.size(2), kw = weight.size(3);
  check_kernel(kernel_h, kernel_w, kh, kw, C, channels_kernel, group);
  const int Ho = (H + 2 * pad_h - (dilation_h * (kernel_h - 1) + 1)) / stride_h + 1;
  const int Wo = (W + 2 * pad_w - (dilation_w * (kernel_w - 1) + 1)) / stride_w + 1;
  at::Tensor x = f32(input), w = f32(weight), off = f32(offset), msk = f32(mask);
  at::Tensor b = with_bias ? f32(bias) : at::Tensor();
  // the reference resizes `output` itself (output.view({B, Co, Ho, Wo}).zero_()); the caller allocated it with that shape
  if (output.numel() != (int64_t)B * Co * Ho * Wo)
    fail("output has %lld elements, expected %lld", (long long)output.numel(), (long long)B * Co * Ho * Wo);
  at::Tensor out = (output.scalar_type() == at::kFloat && output.is_contiguous())
                       ? output
                       : at::empty({B, Co, Ho, Wo}, x.options());
  auto st = at::cuda::getCurrentCUDAStream();
  const size_t need = mrefsr_dcn_workspace_bytes(B, C, H, W, Co, kh, kw, stride_h, stride_w, pad_h, pad_w, dilation_h,
                                                 dilation_w, group, deformable_group, MREFSR_DCN_AUTO, 0);
  Workspace ws = make_workspace(need, x);
  check_rc(mrefsr_modulated_deform_conv_forward(x.data_ptr<float>(), w.data_ptr<float>(),
                                                with_bias ? b.data_ptr<float>() : nullptr, off.data_ptr<float>(),
                                                msk.data_ptr<float>(), out.data_ptr<float>(), B, C, H, W, Co, kh, kw, stride_h,
                                                stride_w, pad_h, pad_w, dilation_h, dilation_w, group, deformable_group,
                                                with_bias ? 1 : 0, MREFSR_DCN_AUTO, ws.ptr, ws.bytes, st.stream()),
           "mrefsr_modulated_deform_conv_forward");
  write_back(output, out);
}

void modulated_deform_conv_backward(at::Tensor input, at::Tensor weight, at::Tensor bias, at::Tensor ones, at::Tensor offset,
                                    at::Tensor mask, at::Tensor columns, at::Tensor grad_input, at::Tensor grad_weight,
                                    at::Tensor grad_bias, at::Tensor grad_offset, at::Tensor grad_mask, at::Tensor grad_output,
                                    int kernel_h, int kernel_w, int stride_h, int stride_w, int pad_h, int pad_w,
                                    int dilation_h, int dilation_w, int group, int deformable_group, const bool with_bias) {
  TORCH_CHECK(input.is_cuda(), "modulated deform conv is not implemented on CPU");
  TORCH_CHECK(input.is_contiguous(), "input tensor has to be contiguous");
  TORCH_CHECK(weight.is_contiguous(), "weight tensor has to be contiguous");
  at::DeviceGuard guard(input.device());
  const int B = input.size(0), C = input.size(1), H = input.size(2), W = input.size(3);
  const int Co = weight.size(0), channels_kernel = weight.size(1), kh = weight.size(2), kw = weight.size(3);
  check_kernel(kernel_h, kernel_w, kh, kw, C, channels_kernel, group);
  at::Tensor x = f32(input), w = f32(weight), off = f32(offset), msk = f32(mask), go = f32(grad_output);
  // gradients: grad_weight / grad_bias ACCUMULATE into the caller's (zero-initialised) tensors as the reference's
  // addmm_ with beta = 1 does (deform_conv_cuda.cpp:659-671); the others are overwritten
  auto grad_buf = [&](at::Tensor& g) {
    return (g.scalar_type() == at::kFloat && g.is_contiguous()) ? g : g.to(at::kFloat).contiguous();
  };
  at::Tensor gi = grad_buf(grad_input), gw = grad_buf(grad_weight), goff = grad_buf(grad_offset), gm = grad_buf(grad_mask);
  at::Tensor gb = with_bias ? grad_buf(grad_bias) : at::Tensor();
  auto st = at::cuda::getCurrentCUDAStream();
  const size_t need = mrefsr_dcn_workspace_bytes(B, C, H, W, Co, kh, kw, stride_h, stride_w, pad_h, pad_w, dilation_h,
                                                 dilation_w, group, deformable_group, MREFSR_DCN_AUTO, 1);
  Workspace ws = make_workspace(need, x);
  check_rc(mrefsr_modulated_deform_conv_backward(
               x.data_ptr<float>(), w.data_ptr<float>(), off.data_ptr<float>(), msk.data_ptr<float>(), go.data_ptr<float>(),
               gi.data_ptr<float>(), gw.data_ptr<float>(), with_bias ? gb.data_ptr<float>() : nullptr,
               goff.data_ptr<float>(), gm.data_ptr<float>(), B, C, H, W, Co, kh, kw, stride_h, stride_w, pad_h, pad_w,
               dilation_h, dilation_w, group, deformable_group, with_bias ? 1 : 0, MREFSR_DCN_AUTO, ws.ptr, ws.bytes,
               st.stream()),
           "mrefsr_modulated_deform_conv_backward");
  write_back(grad_input, gi);
  write_back(grad_weight, gw);
  write_back(grad_offset, goff);
  write_back(grad_mask, gm);
  if (with_bias) write_back(grad_bias, gb);
}

// ---- DCNv1 (deform_conv_ext.cpp:52-105; unused by MRefSR): the same kernels with an implicit all-ones mask.
// Note the reference's (W, H) argument order.
int deform_conv_forward(at::Tensor input, at::Tensor weight, at::Tensor offset, at::Tensor output, at::Tensor columns,
                        at::Tensor ones, int kW, int kH, int dW, int dH, int padW, int padH, int dilationW, int dilationH,
                        int group, int deformable_group, int im2col_step) {
  TORCH_CHECK(input.is_cuda(), "deform conv is not implemented on CPU");
  at::DeviceGuard guard(input.device());
  at::Tensor x = f32(input), w = f32(weight), off = f32(offset);
  const int B = x.size(0), C = x.size(1), H = x.size(2), W = x.size(3), Co = w.size(0);
  const int Ho = (H + 2 * padH - (dilationH * (kH - 1) + 1)) / dH + 1, Wo = (W + 2 * padW - (dilationW * (kW - 1) + 1)) / dW + 1;
  at::Tensor out = at::empty({B, Co, Ho, Wo}, x.options());
  auto st = at::cuda::getCurrentCUDAStream();
  const size_t need = mrefsr_dcn_workspace_bytes(B, C, H, W, Co, kH, kW, dH, dW, padH, padW, dilationH, dilationW, group,
                                                 deformable_group, MREFSR_DCN_AUTO, 0);
  Workspace ws = make_workspace(need, x);
  check_rc(mrefsr_modulated_deform_conv_forward(x.data_ptr<float>(), w.data_ptr<float>(), nullptr, off.data_ptr<float>(),
                                                nullptr, out.data_ptr<float>(), B, C, H, W, Co, kH, kW, dH, dW, padH, padW,
                                                dilationH, dilationW, group, deformable_group, 0, MREFSR_DCN_AUTO, ws.ptr,
                                                ws.bytes, st.stream()),
           "mrefsr_modulated_deform_conv_forward (DCNv1)");
  output.resize_({B, Co, Ho, Wo});
  output.copy_(out);
  return 1;
}

int deform_conv_backward_input(at::Tensor input, at::Tensor offset, at::Tensor gradOutput, at::Tensor gradInput,
                               at::Tensor gradOffset, at::Tensor weight, at::Tensor columns, int kW, int kH, int dW, int dH,
                               int padW, int padH, int dilationW, int dilationH, int group, int deformable_group,
                               int im2col_step) {
  TORCH_CHECK(input.is_cuda(), "deform conv is not implemented on CPU");
  at::DeviceGuard guard(input.device());
  at::Tensor x = f32(input), w = f32(weight), off = f32(offset), go = f32(gradOutput);
  const int B = x.size(0), C = x.size(1), H = x.size(2), W = x.size(3), Co = w.size(0);
  at::Tensor gi = at::zeros_like(x), goff = at::zeros_like(off);
  auto st = at::cuda::getCurrentCUDAStream();
  const size_t need = mrefsr_dcn_workspace_bytes(B, C, H, W, Co, kH, kW, dH, dW, padH, padW, dilationH, dilationW, group,
                                                 deformable_group, MREFSR_DCN_AUTO, 1);
  Workspace ws = make_workspace(need, x);
  check_rc(mrefsr_modulated_deform_conv_backward(x.data_ptr<float>(), w.data_ptr<float>(), off.data_ptr<float>(), nullptr,
                                                 go.data_ptr<float>(), gi.data_ptr<float>(), nullptr, nullptr,
                                                 goff.data_ptr<float>(), nullptr, B, C, H, W, Co, kH, kW, dH, dW, padH, padW,
                                                 dilationH, dilationW, group, deformable_group, 0, MREFSR_DCN_AUTO, ws.ptr,
                                                 ws.bytes, st.stream()),
           "mrefsr_modulated_deform_conv_backward (DCNv1 input)");
  gradInput.resize_as_(gi).copy_(gi);
  gradOffset.resize_as_(goff).copy_(goff);
  return 1;
}

int deform_conv_backward_parameters(at::Tensor input, at::Tensor offset, at::Tensor gradOutput, at::Tensor gradWeight,
                                    at::Tensor columns, at::Tensor ones, int kW, int kH, int dW, int dH, int padW, int padH,
                                    int dilationW, int dilationH, int group, int deformable_group, float scale,
                                    int im2col_step) {
  TORCH_CHECK(input.is_cuda(), "deform conv is not implemented on CPU");
  at::DeviceGuard guard(input.device());
  at::Tensor x = f32(input), off = f32(offset), go = f32(gradOutput);
  const int B = x.size(0), C = x.size(1), H = x.size(2), W = x.size(3), Co = gradWeight.size(0);
  at::Tensor w = at::zeros({Co, C / group, kH, kW}, x.options());     // the weights themselves do not enter grad_weight
  at::Tensor gw = at::zeros_like(w);
  auto st = at::cuda::getCurrentCUDAStream();
  const size_t need = mrefsr_dcn_workspace_bytes(B, C, H, W, Co, kH, kW, dH, dW, padH, padW, dilationH, dilationW, group,
                                                 deformable_group, MREFSR_DCN_AUTO, 1);
  Workspace ws = make_workspace(need, x);
  check_rc(mrefsr_modulated_deform_conv_backward(x.data_ptr<float>(), w.data_ptr<float>(), off.data_ptr<float>(), nullptr,
                                                 go.data_ptr<float>(), nullptr, gw.data_ptr<float>(), nullptr, nullptr,
                                                 nullptr, B, C, H, W, Co, kH, kW, dH, dW, padH, padW, dilationH, dilationW,
                                                 group, deformable_group, 0, MREFSR_DCN_AUTO, ws.ptr, ws.bytes, st.stream()),
           "mrefsr_modulated_deform_conv_backward (DCNv1 parameters)");
  gradWeight.add_(gw.view_as(gradWeight).to(gradWeight.scalar_type()), scale);     // accumulates, like the reference
  return 1;
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("deform_conv_forward", &deform_conv_forward, "deform forward");
  m.def("deform_conv_backward_input", &deform_conv_backward_input, "deform_conv_backward_input");
  m.def("deform_conv_backward_parameters", &deform_conv_backward_parameters, "deform_conv_backward_parameters");
  m.def("modulated_deform_conv_forward", &modulated_deform_conv_forward, "modulated deform conv forward");
  m.def("modulated_deform_conv_backward", &modulated_deform_conv_backward, "modulated deform conv backward");
}
