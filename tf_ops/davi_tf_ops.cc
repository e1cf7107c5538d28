// TensorFlow custom-op shim over the davi C ABI (include/davi.h).
//
// STATUS: TensorFlow (headers and runtime) is absent from the build container
// and from the GPU boxes, so this file cannot be linked or run here; every line
// of arithmetic lives behind the C ABI and is exercised through the ctypes
// binding (deep_attentive_vi_b200/_lib.py).  tests/test_tf_shim.py compiles this
// file with -fsyntax-only against a stand-in for the op-kernel headers
// (tf_ops/tf_stub/) so that every C-ABI call site below is type-checked against
// include/davi.h.  Where TensorFlow exists, build it with
//
//   TF_CFLAGS=$(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_compile_flags()))')
//   TF_LFLAGS=$(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_link_flags()))')
//   g++ -std=c++17 -shared -fPIC tf_ops/davi_tf_ops.cc -o tf_ops/libdavi_tf.so \
//       -Iinclude $TF_CFLAGS $TF_LFLAGS -DGOOGLE_CUDA=1 \
//       -Ldeep_attentive_vi_b200 -ldavi -Wl,-rpath,'$ORIGIN/../deep_attentive_vi_b200'
//
// and load it with tf.load_op_library (tf_ops/davi_tf.py registers the
// gradients and subclasses the reference layers).
//
// Each op forwards raw device pointers + the op's CUDA stream to ONE C-ABI
// entry point; outputs and workspaces are allocated by TensorFlow
// (allocate_output / allocate_temp), errors become errors::Internal with the
// text of davi_last_error().  There is no CPU kernel registration: placing one
// of these ops on a CPU device fails at graph construction.
#define EIGEN_USE_GPU
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"
#include "tensorflow/core/util/gpu_kernel_helper.h"

#include "davi.h"

namespace tf = tensorflow;
using tf::OpKernel;
using tf::OpKernelConstruction;
using tf::OpKernelContext;
using tf::Tensor;
using tf::TensorShape;

namespace {

davi_stream_t StreamOf(OpKernelContext* ctx) {
  return reinterpret_cast<davi_stream_t>(ctx->eigen_gpu_device().stream());
}

tf::Status Check(int rc, const char* what) {
  if (rc == DAVI_OK) return tf::Status::OK();
  char buf[512];
  davi_last_error(buf, sizeof(buf));
  return tf::errors::Internal(what, " failed (", rc, "): ", buf);
}

const float* P(const Tensor& t) { return t.flat<float>().data(); }
float* P(Tensor* t) { return t->flat<float>().data(); }

}  // namespace

// ---------------------------------------------------------------------------
// DaviDwAttn: DepthWiseAttention.apply (src/layers/attention.py:135-164)
//   query [B,W,H,d], keys [B,n,W,H,d], values [B,n,W,H,C], scale [] -> [B,W,H,C]
// ---------------------------------------------------------------------------
REGISTER_OP("DaviDwAttn")
    .Input("query: float")
    .Input("keys: float")
    .Input("values: float")
    .Input("scale: float")
    .Output("out: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      tf::shape_inference::ShapeHandle v, out;
      TF_RETURN_IF_ERROR(c->WithRank(c->input(2), 5, &v));
      out = c->MakeShape({c->Dim(v, 0), c->Dim(v, 2), c->Dim(v, 3), c->Dim(v, 4)});
      c->set_output(0, out);
      return tf::Status::OK();
    });

class DaviDwAttnOp : public OpKernel {
 public:
  explicit DaviDwAttnOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor &q = ctx->input(0), &k = ctx->input(1), &v = ctx->input(2), &s = ctx->input(3);
    OP_REQUIRES(ctx, k.dims() == 5 && v.dims() == 5 && q.dims() == 4,
                tf::errors::InvalidArgument("DaviDwAttn: expected ranks 4/5/5"));
    const int B = k.dim_size(0), n = k.dim_size(1), W = k.dim_size(2), H = k.dim_size(3);
    const int d = k.dim_size(4), C = v.dim_size(4), Pn = W * H;
    Tensor* out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({B, W, H, C}), &out));
    OP_REQUIRES_OK(ctx, Check(davi_dw_attn_fwd(P(q), P(k), P(v), P(out), nullptr, B, n, Pn, d, C,
                                               d, (int64_t)n * Pn * d, (int64_t)Pn * d, d,
                                               (int64_t)n * Pn * C, (int64_t)Pn * C, C, C, P(s),
                                               StreamOf(ctx)),
                              "davi_dw_attn_fwd"));
  }
};
REGISTER_KERNEL_BUILDER(Name("DaviDwAttn").Device(tf::DEVICE_GPU), DaviDwAttnOp);

REGISTER_OP("DaviDwAttnGrad")
    .Input("query: float")
    .Input("keys: float")
    .Input("values: float")
    .Input("scale: float")
    .Input("dout: float")
    .Output("dquery: float")
    .Output("dkeys: float")
    .Output("dvalues: float")
    .Output("dscale: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      for (int i = 0; i < 4; ++i) c->set_output(i, c->input(i));
      return tf::Status::OK();
    });

class DaviDwAttnGradOp : public OpKernel {
 public:
  explicit DaviDwAttnGradOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor &q = ctx->input(0), &k = ctx->input(1), &v = ctx->input(2), &s = ctx->input(3);
    const Tensor& go = ctx->input(4);
    const int B = k.dim_size(0), n = k.dim_size(1), W = k.dim_size(2), H = k.dim_size(3);
    const int d = k.dim_size(4), C = v.dim_size(4), Pn = W * H;
    Tensor *dq, *dk, *dv, *ds;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, q.shape(), &dq));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, k.shape(), &dk));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, v.shape(), &dv));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(3, TensorShape({}), &ds));
    OP_REQUIRES_OK(ctx, Check(davi_dw_attn_bwd(P(q), P(k), P(v), P(go), P(dq), P(dk), P(dv), P(ds),
                                               B, n, Pn, d, C, d, (int64_t)n * Pn * d,
                                               (int64_t)Pn * d, d, (int64_t)n * Pn * C,
                                               (int64_t)Pn * C, C, C, P(s), StreamOf(ctx)),
                              "davi_dw_attn_bwd"));
  }
};
REGISTER_KERNEL_BUILDER(Name("DaviDwAttnGrad").Device(tf::DEVICE_GPU), DaviDwAttnGradOp);

// ---------------------------------------------------------------------------
// DaviNonlocalAttn: core of NonlocalResNetBlock.call (attention.py:318-389)
//   q,k [B,W,H,d]; v,x [B,W,H,C]; scale, gamma [] -> y [B,W,H,C], lse [B,W*H]
// ---------------------------------------------------------------------------
REGISTER_OP("DaviNonlocalAttn")
    .Input("q: float").Input("k: float").Input("v: float").Input("x: float")
    .Input("scale: float").Input("gamma: float")
    .Output("y: float").Output("lse: float").Output("o: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      c->set_output(0, c->input(3));
      c->set_output(1, c->UnknownShapeOfRank(2));
      c->set_output(2, c->input(3));
      return tf::Status::OK();
    });

class DaviNonlocalAttnOp : public OpKernel {
 public:
  explicit DaviNonlocalAttnOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor &q = ctx->input(0), &k = ctx->input(1), &v = ctx->input(2), &x = ctx->input(3);
    const int B = x.dim_size(0), N = x.dim_size(1) * x.dim_size(2);
    const int d = q.dim_size(3), C = x.dim_size(3);
    Tensor *y, *lse, *o;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, x.shape(), &y));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, TensorShape({B, N}), &lse));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, x.shape(), &o));   // copy of O for the gradient op
    OP_REQUIRES_OK(ctx, Check(davi_nonlocal_attn_fwd(P(q), P(k), P(v), P(x), P(y), P(lse), P(o), B, N, d, C,
                                                     d, d, C, C, C, C, P(ctx->input(4)), P(ctx->input(5)),
                                                     StreamOf(ctx)),
                              "davi_nonlocal_attn_fwd"));
  }
};
REGISTER_KERNEL_BUILDER(Name("DaviNonlocalAttn").Device(tf::DEVICE_GPU), DaviNonlocalAttnOp);

REGISTER_OP("DaviNonlocalAttnGrad")
    .Input("q: float").Input("k: float").Input("v: float").Input("lse: float")
    .Input("scale: float").Input("gamma: float").Input("dy: float").Input("o: float")
    .Output("dq: float").Output("dk: float").Output("dv: float").Output("dgamma_dscale: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      for (int i = 0; i < 3; ++i) c->set_output(i, c->input(i));
      c->set_output(3, c->Vector(2));
      return tf::Status::OK();
    });

class DaviNonlocalAttnGradOp : public OpKernel {
 public:
  explicit DaviNonlocalAttnGradOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor &q = ctx->input(0), &k = ctx->input(1), &v = ctx->input(2), &lse = ctx->input(3);
    const Tensor &dy = ctx->input(6), &o = ctx->input(7);
    const int B = v.dim_size(0), N = v.dim_size(1) * v.dim_size(2);
    const int d = q.dim_size(3), C = v.dim_size(3);
    Tensor *dq, *dk, *dv, *dgs, ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, q.shape(), &dq));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, k.shape(), &dk));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, v.shape(), &dv));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(3, TensorShape({2}), &dgs));
    const size_t wsb = davi_nonlocal_attn_bwd_workspace_bytes(B, N, d, C);
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_FLOAT, TensorShape({(tf::int64)((wsb + 3) / 4)}), &ws));
    OP_REQUIRES_OK(ctx, Check(davi_nonlocal_attn_bwd(P(q), P(k), P(v), P(lse), P(o), P(dy), P(dq), P(dk), P(dv),
                                                     P(dgs), B, N, d, C, d, d, C, C, C, P(ctx->input(4)),
                                                     P(ctx->input(5)), P(&ws), wsb, StreamOf(ctx)),
                              "davi_nonlocal_attn_bwd"));
  }
};
REGISTER_KERNEL_BUILDER(Name("DaviNonlocalAttnGrad").Device(tf::DEVICE_GPU), DaviNonlocalAttnGradOp);

// ---------------------------------------------------------------------------
// DaviGaussKl: GaussianDiag residual posterior sample + KL
//   (stat_tools/gaussian_diag.py:121-240, model/variational_layer.py:511-542)
// ---------------------------------------------------------------------------
REGISTER_OP("DaviGaussKl")
    .Input("post: float").Input("prior: float").Input("eps: float")
    .Input("noise_post: float").Input("noise_prior: float")
    .Attr("post_lo: float").Attr("post_hi: float").Attr("prior_lo: float").Attr("prior_hi: float")
    .Attr("noise_std_post: float = 0.0").Attr("noise_std_prior: float = 0.0")
    .Attr("has_prior: bool = true")
    .Output("xi: float").Output("kl: float").Output("logq: float").Output("logp: float")
    .Output("grad: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      c->set_output(0, c->input(2));
      for (int i = 1; i < 4; ++i) c->set_output(i, c->Vector(c->Dim(c->input(2), 0)));
      c->set_output(4, c->UnknownShapeOfRank(4));
      return tf::Status::OK();
    });

// noise_post / noise_prior: the N(0,1) draws of the train-time GaussianNoise on the log-scales
// (gaussian_diag.py:165-166), same shape as eps; an EMPTY tensor (0 elements) = no noise.
// grad [B,W,H,6z]: the whole gradient, emitted by the forward sweep (include/davi.h, davi_gauss_kl_fwd).
class DaviGaussKlOp : public OpKernel {
 public:
  explicit DaviGaussKlOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("post_lo", &qlo_));
    OP_REQUIRES_OK(c, c->GetAttr("post_hi", &qhi_));
    OP_REQUIRES_OK(c, c->GetAttr("prior_lo", &plo_));
    OP_REQUIRES_OK(c, c->GetAttr("prior_hi", &phi_));
    OP_REQUIRES_OK(c, c->GetAttr("noise_std_post", &nq_));
    OP_REQUIRES_OK(c, c->GetAttr("noise_std_prior", &np_));
    OP_REQUIRES_OK(c, c->GetAttr("has_prior", &has_prior_));
  }
  void Compute(OpKernelContext* ctx) override {
    const Tensor &post = ctx->input(0), &prior = ctx->input(1), &eps = ctx->input(2);
    const Tensor &n1 = ctx->input(3), &n2 = ctx->input(4);
    const int B = eps.dim_size(0), W = eps.dim_size(1), H = eps.dim_size(2), z = eps.dim_size(3);
    const int Pn = W * H;
    Tensor *xi, *kl, *lq, *lp, *grad;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, eps.shape(), &xi));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, TensorShape({B}), &kl));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, TensorShape({B}), &lq));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(3, TensorShape({B}), &lp));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(4, TensorShape({B, W, H, 6 * z}), &grad));
    OP_REQUIRES_OK(ctx, Check(davi_gauss_kl_fwd(P(post), has_prior_ ? P(prior) : nullptr, P(eps),
                                                n1.NumElements() ? P(n1) : nullptr,
                                                has_prior_ && n2.NumElements() ? P(n2) : nullptr,
                                                P(xi), P(kl), P(lq), P(lp), B, Pn, z, qlo_, qhi_,
                                                plo_, phi_, nq_, np_, P(grad), StreamOf(ctx)),
                              "davi_gauss_kl_fwd"));
  }
 private:
  float qlo_, qhi_, plo_, phi_, nq_, np_;
  bool has_prior_;
};
REGISTER_KERNEL_BUILDER(Name("DaviGaussKl").Device(tf::DEVICE_GPU), DaviGaussKlOp);

REGISTER_OP("DaviGaussKlGrad")
    .Input("grad: float").Input("dxi: float").Input("dkl: float")
    .Attr("has_prior: bool = true")
    .Output("dpost: float").Output("dprior: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      c->set_output(0, c->UnknownShapeOfRank(4));
      c->set_output(1, c->UnknownShapeOfRank(4));
      return tf::Status::OK();
    });

class DaviGaussKlGradOp : public OpKernel {
 public:
  explicit DaviGaussKlGradOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("has_prior", &has_prior_));
  }
  void Compute(OpKernelContext* ctx) override {
    const Tensor &grad = ctx->input(0), &dxi = ctx->input(1), &dkl = ctx->input(2);
    const int B = grad.dim_size(0), W = grad.dim_size(1), H = grad.dim_size(2), z = grad.dim_size(3) / 6;
    Tensor *dpost, *dprior;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({B, W, H, 2 * z}), &dpost));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, TensorShape({B, W, H, 2 * z}), &dprior));
    OP_REQUIRES_OK(ctx, Check(davi_gauss_kl_bwd_saved(P(grad), P(dxi), P(dkl), P(dpost),
                                                      has_prior_ ? P(dprior) : nullptr, B, W * H, z,
                                                      StreamOf(ctx)),
                              "davi_gauss_kl_bwd_saved"));
  }
 private:
  bool has_prior_;
};
REGISTER_KERNEL_BUILDER(Name("DaviGaussKlGrad").Device(tf::DEVICE_GPU), DaviGaussKlGradOp);

// ---------------------------------------------------------------------------
// DaviLnGelu / DaviLnGeluRes: the LayerNormalization / gelu / gamma-residual glue around the depth-wise
// attention (model/variational_layer.py:393-412,501-505; model/vae.py:346-348)
//   mode 0: y = LN(x)    mode 1: y = x + gelu(LN(x))    Res: y = base + res_gamma * (x + gelu(LN(x)))
// ---------------------------------------------------------------------------
REGISTER_OP("DaviLnGelu")
    .Input("x: float").Input("gamma: float").Input("beta: float")
    .Attr("mode: int = 0").Attr("epsilon: float = 1e-5")
    .Output("y: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      c->set_output(0, c->input(0));
      return tf::Status::OK();
    });

class DaviLnGeluOp : public OpKernel {
 public:
  explicit DaviLnGeluOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("mode", &mode_));
    OP_REQUIRES_OK(c, c->GetAttr("epsilon", &eps_));
  }
  void Compute(OpKernelContext* ctx) override {
    const Tensor &x = ctx->input(0), &g = ctx->input(1), &b = ctx->input(2);
    const int C = x.dim_size(x.dims() - 1);
    const tf::int64 rows = x.NumElements() / C;
    Tensor* y;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, x.shape(), &y));
    OP_REQUIRES_OK(ctx, Check(davi_ln_gelu_fwd(P(x), P(g), P(b), P(y), rows, C, C, C, eps_, mode_, StreamOf(ctx)),
                              "davi_ln_gelu_fwd"));
  }
 private:
  int mode_;
  float eps_;
};
REGISTER_KERNEL_BUILDER(Name("DaviLnGelu").Device(tf::DEVICE_GPU), DaviLnGeluOp);

REGISTER_OP("DaviLnGeluGrad")
    .Input("x: float").Input("gamma: float").Input("beta: float").Input("dy: float")
    .Attr("mode: int = 0").Attr("epsilon: float = 1e-5")
    .Output("dx: float").Output("dgamma: float").Output("dbeta: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      for (int i = 0; i < 3; ++i) c->set_output(i, c->input(i));
      return tf::Status::OK();
    });

class DaviLnGeluGradOp : public OpKernel {
 public:
  explicit DaviLnGeluGradOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("mode", &mode_));
    OP_REQUIRES_OK(c, c->GetAttr("epsilon", &eps_));
  }
  void Compute(OpKernelContext* ctx) override {
    const Tensor &x = ctx->input(0), &g = ctx->input(1), &b = ctx->input(2), &dy = ctx->input(3);
    const int C = x.dim_size(x.dims() - 1);
    const tf::int64 rows = x.NumElements() / C;
    Tensor *dx, *dg, *db, ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, x.shape(), &dx));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, g.shape(), &dg));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, b.shape(), &db));
    const size_t wsb = davi_ln_gelu_bwd_workspace_bytes(rows, C);
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_FLOAT, TensorShape({(tf::int64)((wsb + 3) / 4)}), &ws));
    OP_REQUIRES_OK(ctx, Check(davi_ln_gelu_bwd(P(x), P(g), P(b), P(dy), P(dx), P(dg), P(db), rows, C, C, C, C,
                                               eps_, mode_, P(&ws), wsb, StreamOf(ctx)),
                              "davi_ln_gelu_bwd"));
  }
 private:
  int mode_;
  float eps_;
};
REGISTER_KERNEL_BUILDER(Name("DaviLnGeluGrad").Device(tf::DEVICE_GPU), DaviLnGeluGradOp);

REGISTER_OP("DaviLnGeluRes")
    .Input("x: float").Input("gamma: float").Input("beta: float").Input("base: float").Input("res_gamma: float")
    .Attr("epsilon: float = 1e-5")
    .Output("y: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      c->set_output(0, c->input(0));
      return tf::Status::OK();
    });

class DaviLnGeluResOp : public OpKernel {
 public:
  explicit DaviLnGeluResOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("epsilon", &eps_));
  }
  void Compute(OpKernelContext* ctx) override {
    const Tensor &x = ctx->input(0), &g = ctx->input(1), &b = ctx->input(2), &base = ctx->input(3);
    const int C = x.dim_size(x.dims() - 1);
    const tf::int64 rows = x.NumElements() / C;
    Tensor* y;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, x.shape(), &y));
    OP_REQUIRES_OK(ctx, Check(davi_ln_gelu_res_fwd(P(x), P(g), P(b), P(base), P(ctx->input(4)), P(y), rows, C,
                                                   C, C, C, eps_, StreamOf(ctx)),
                              "davi_ln_gelu_res_fwd"));
  }
 private:
  float eps_;
};
REGISTER_KERNEL_BUILDER(Name("DaviLnGeluRes").Device(tf::DEVICE_GPU), DaviLnGeluResOp);

REGISTER_OP("DaviLnGeluResGrad")
    .Input("x: float").Input("gamma: float").Input("beta: float").Input("res_gamma: float").Input("dy: float")
    .Attr("epsilon: float = 1e-5")
    .Output("dx: float").Output("dgamma: float").Output("dbeta: float").Output("dres_gamma: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      for (int i = 0; i < 4; ++i) c->set_output(i, c->input(i));
      return tf::Status::OK();
    });

class DaviLnGeluResGradOp : public OpKernel {
 public:
  explicit DaviLnGeluResGradOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("epsilon", &eps_));
  }
  void Compute(OpKernelContext* ctx) override {
    const Tensor &x = ctx->input(0), &g = ctx->input(1), &b = ctx->input(2), &rg = ctx->input(3);
    const Tensor& dy = ctx->input(4);
    const int C = x.dim_size(x.dims() - 1);
    const tf::int64 rows = x.NumElements() / C;
    Tensor *dx, *dg, *db, *drg, ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, x.shape(), &dx));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, g.shape(), &dg));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, b.shape(), &db));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(3, TensorShape({}), &drg));
    const size_t wsb = davi_ln_gelu_bwd_workspace_bytes(rows, C);
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_FLOAT, TensorShape({(tf::int64)((wsb + 3) / 4)}), &ws));
    OP_REQUIRES_OK(ctx, Check(davi_ln_gelu_res_bwd(P(x), P(g), P(b), P(rg), P(dy), P(dx), P(dg), P(db), P(drg),
                                                   rows, C, C, C, C, eps_, P(&ws), wsb, StreamOf(ctx)),
                              "davi_ln_gelu_res_bwd"));
  }
 private:
  float eps_;
};
REGISTER_KERNEL_BUILDER(Name("DaviLnGeluResGrad").Device(tf::DEVICE_GPU), DaviLnGeluResGradOp);

// ---------------------------------------------------------------------------
// DaviDlmNll / DaviBernoulliNll: data negative log-likelihood with the
// per-image gradient emitted by the same launch
//   (stat_tools/discretized_logistic_mixture.py:161-211,311-412;
//    stat_tools/bernoulli.py:155-179; model/data_layer.py:179-181)
// ---------------------------------------------------------------------------
REGISTER_OP("DaviDlmNll")
    .Input("params: float").Input("x: float")
    .Attr("log_scale_low_bound: float = -7.0")
    .Output("nll: float").Output("dparams: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      c->set_output(0, c->Vector(c->Dim(c->input(0), 0)));
      c->set_output(1, c->input(0));
      return tf::Status::OK();
    });

class DaviDlmNllOp : public OpKernel {
 public:
  explicit DaviDlmNllOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("log_scale_low_bound", &lo_));
  }
  void Compute(OpKernelContext* ctx) override {
    const Tensor &p = ctx->input(0), &x = ctx->input(1);
    const int B = p.dim_size(0), Pn = p.dim_size(1) * p.dim_size(2);
    Tensor *nll, *dp;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({B}), &nll));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, p.shape(), &dp));
    OP_REQUIRES_OK(ctx, Check(davi_dlm_nll_fwd(P(p), P(x), P(nll), P(dp), B, Pn, lo_, StreamOf(ctx)),
                              "davi_dlm_nll_fwd"));
  }
 private:
  float lo_;
};
REGISTER_KERNEL_BUILDER(Name("DaviDlmNll").Device(tf::DEVICE_GPU), DaviDlmNllOp);

REGISTER_OP("DaviBernoulliNll")
    .Input("logits: float").Input("x: float")
    .Attr("crop: int = 0")
    .Output("nll: float").Output("dlogits: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      c->set_output(0, c->Vector(c->Dim(c->input(0), 0)));
      c->set_output(1, c->input(0));
      return tf::Status::OK();
    });

class DaviBernoulliNllOp : public OpKernel {
 public:
  explicit DaviBernoulliNllOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("crop", &crop_));
  }
  void Compute(OpKernelContext* ctx) override {
    const Tensor &l = ctx->input(0), &x = ctx->input(1);
    const int B = l.dim_size(0), H = l.dim_size(1), W = l.dim_size(2);
    Tensor *nll, *dl;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({B}), &nll));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, l.shape(), &dl));
    OP_REQUIRES_OK(ctx, Check(davi_bernoulli_nll_fwd(P(l), P(x), P(nll), P(dl), B, H, W, crop_,
                                                     StreamOf(ctx)),
                              "davi_bernoulli_nll_fwd"));
  }
 private:
  int crop_;
};
REGISTER_KERNEL_BUILDER(Name("DaviBernoulliNll").Device(tf::DEVICE_GPU), DaviBernoulliNllOp);

// ---------------------------------------------------------------------------
// DaviSplitTf32: three-part TF32 operand split for fp32-grade Conv2D on the tensor cores
// (row f1; src/layers/resnet.py:589-601, src/layers/weight_norm.py:126-135)
//   x [..., C] -> layout 0: [..., 3C] ; layout 1: [3 * dim0, ..., C]
// ---------------------------------------------------------------------------
REGISTER_OP("DaviSplitTf32")
    .Input("x: float")
    .Attr("layout: int = 0").Attr("pattern: int = 0")
    .Output("parts: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      c->set_output(0, c->UnknownShape());
      return tf::Status::OK();
    });

class DaviSplitTf32Op : public OpKernel {
 public:
  explicit DaviSplitTf32Op(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("layout", &layout_));
    OP_REQUIRES_OK(c, c->GetAttr("pattern", &pattern_));
  }
  void Compute(OpKernelContext* ctx) override {
    const Tensor& x = ctx->input(0);
    const int nd = x.dims();
    OP_REQUIRES(ctx, nd >= 1, tf::errors::InvalidArgument("DaviSplitTf32: rank >= 1"));
    const int C = x.dim_size(nd - 1);
    const tf::int64 rows = x.NumElements() / C;
    TensorShape shape = x.shape();
    if (layout_ == 0) shape.set_dim(nd - 1, 3 * C); else shape.set_dim(0, 3 * x.dim_size(0));
    Tensor* out;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, shape, &out));
    OP_REQUIRES_OK(ctx, Check(davi_split_tf32(P(x), P(out), rows, C, layout_, pattern_, StreamOf(ctx)),
                              "davi_split_tf32"));
  }
 private:
  int layout_, pattern_;
};
REGISTER_KERNEL_BUILDER(Name("DaviSplitTf32").Device(tf::DEVICE_GPU), DaviSplitTf32Op);

// ---------------------------------------------------------------------------
// DaviConv2d / DaviConv2dInputGrad / DaviConv2dFilterGrad: the Conv2D of the ResNet cells on the tcgen05 kernels
// (row f1; src/layers/resnet.py:589-601,613-633, src/layers/weight_norm.py:126-135)
//   x [B,H,W,Cin], kernel_ohwi [Cout,kh,kw,Cin] (tf.transpose(kernel_hwio, [3,0,1,2])), bias [Cout] -> [B,H,W,Cout]
// stride 1, padding 'same'.  The stream-K workspace is a persistent buffer of the kernel object, zero-filled once.
// ---------------------------------------------------------------------------
namespace {
class ConvWorkspace {
 public:
  tf::Status Get(OpKernelContext* ctx, void** ptr, size_t* bytes) {
    *bytes = davi_conv2d_workspace_bytes();
    if (!ready_) {
      TF_RETURN_IF_ERROR(ctx->allocate_temp(tf::DT_FLOAT, TensorShape({(tf::int64)(*bytes / 4)}), &ws_));
      TF_RETURN_IF_ERROR(Check(davi_fill_zero(P(&ws_), *bytes / 4, StreamOf(ctx)), "davi_fill_zero"));
      ready_ = true;
    }
    *ptr = P(&ws_);
    return tf::Status::OK();
  }
 private:
  Tensor ws_;            // kept alive by the op kernel: the hand-shake flags inside must survive between calls
  bool ready_ = false;
};
}  // namespace

REGISTER_OP("DaviConv2d")
    .Input("x: float").Input("kernel_ohwi: float").Input("bias: float")
    .Output("y: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      c->set_output(0, c->UnknownShape());
      return tf::Status::OK();
    });

class DaviConv2dOp : public OpKernel {
 public:
  explicit DaviConv2dOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor& x = ctx->input(0);
    const Tensor& w = ctx->input(1);
    const Tensor& b = ctx->input(2);
    OP_REQUIRES(ctx, x.dims() == 4 && w.dims() == 4 && w.dim_size(3) == x.dim_size(3),
                tf::errors::InvalidArgument("DaviConv2d: x [B,H,W,Cin], kernel [Cout,kh,kw,Cin]"));
    const int B = x.dim_size(0), H = x.dim_size(1), W = x.dim_size(2), Cin = x.dim_size(3);
    const int Cout = w.dim_size(0), kh = w.dim_size(1), kw = w.dim_size(2);
    Tensor* y;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({B, H, W, Cout}), &y));
    void* ws; size_t wsb;
    OP_REQUIRES_OK(ctx, ws_.Get(ctx, &ws, &wsb));
    OP_REQUIRES_OK(ctx, Check(davi_conv2d_fwd(P(x), P(w), b.NumElements() ? P(b) : nullptr, P(y), B, H, W, Cin, Cout,
                                              kh, kw, Cin, Cout, ws, wsb, StreamOf(ctx)), "davi_conv2d_fwd"));
  }
 private:
  ConvWorkspace ws_;
};
REGISTER_KERNEL_BUILDER(Name("DaviConv2d").Device(tf::DEVICE_GPU), DaviConv2dOp);

// DaviTf32Lo / DaviConv2dPre: kernels that live across many calls (once per optimiser step under WeightNorm) carry the lo
// half of their 3xTF32 split with them: kernel_lo = DaviTf32Lo(kernel_ohwi) once per step, then every DaviConv2dPre of
// the step loads it by TMA instead of deriving it from each staged weight tile (same bits out).
REGISTER_OP("DaviTf32Lo")
    .Input("x: float")
    .Output("lo: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      c->set_output(0, c->input(0));
      return tf::Status::OK();
    });

class DaviTf32LoOp : public OpKernel {
 public:
  explicit DaviTf32LoOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor& x = ctx->input(0);
    Tensor* lo;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, x.shape(), &lo));
    OP_REQUIRES_OK(ctx, Check(davi_tf32_lo(P(x), P(lo), (int64_t)x.NumElements(), StreamOf(ctx)), "davi_tf32_lo"));
  }
};
REGISTER_KERNEL_BUILDER(Name("DaviTf32Lo").Device(tf::DEVICE_GPU), DaviTf32LoOp);

REGISTER_OP("DaviConv2dPre")
    .Input("x: float").Input("kernel_ohwi: float").Input("kernel_lo: float").Input("bias: float")
    .Attr("swish_input: bool = false")
    .Output("y: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      c->set_output(0, c->UnknownShape());
      return tf::Status::OK();
    });

class DaviConv2dPreOp : public OpKernel {
 public:
  explicit DaviConv2dPreOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("swish_input", &swish_));
  }
  void Compute(OpKernelContext* ctx) override {
    const Tensor& x = ctx->input(0);
    const Tensor& w = ctx->input(1);
    const Tensor& wl = ctx->input(2);
    const Tensor& b = ctx->input(3);
    OP_REQUIRES(ctx, x.dims() == 4 && w.dims() == 4 && w.dim_size(3) == x.dim_size(3) && wl.dims() == 4 &&
                         wl.NumElements() == w.NumElements() && wl.dim_size(0) == w.dim_size(0) && wl.dim_size(3) == w.dim_size(3),
                tf::errors::InvalidArgument("DaviConv2dPre: x [B,H,W,Cin], kernel and kernel_lo [Cout,kh,kw,Cin]"));
    const int B = x.dim_size(0), H = x.dim_size(1), W = x.dim_size(2), Cin = x.dim_size(3);
    const int Cout = w.dim_size(0), kh = w.dim_size(1), kw = w.dim_size(2);
    Tensor* y;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({B, H, W, Cout}), &y));
    void* ws; size_t wsb;
    OP_REQUIRES_OK(ctx, ws_.Get(ctx, &ws, &wsb));
    OP_REQUIRES_OK(ctx, Check(davi_conv2d_fwd_pre(P(x), P(w), P(wl), b.NumElements() ? P(b) : nullptr, P(y), B, H, W, Cin,
                                                  Cout, kh, kw, Cin, Cout, swish_ ? 1 : 0, ws, wsb, StreamOf(ctx)),
                              "davi_conv2d_fwd_pre"));
  }
 private:
  ConvWorkspace ws_;
  bool swish_ = false;
};
REGISTER_KERNEL_BUILDER(Name("DaviConv2dPre").Device(tf::DEVICE_GPU), DaviConv2dPreOp);

REGISTER_OP("DaviConv2dInputGrad")
    .Input("dy: float").Input("kernel_ohwi: float")
    .Output("dx: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      c->set_output(0, c->UnknownShape());
      return tf::Status::OK();
    });

class DaviConv2dInputGradOp : public OpKernel {
 public:
  explicit DaviConv2dInputGradOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor& dy = ctx->input(0);
    const Tensor& w = ctx->input(1);
    const int B = dy.dim_size(0), H = dy.dim_size(1), W = dy.dim_size(2), Cout = dy.dim_size(3);
    const int kh = w.dim_size(1), kw = w.dim_size(2), Cin = w.dim_size(3);
    OP_REQUIRES(ctx, w.dim_size(0) == Cout, tf::errors::InvalidArgument("DaviConv2dInputGrad: channel mismatch"));
    Tensor wt;
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_FLOAT, TensorShape({Cin, kh, kw, Cout}), &wt));
    OP_REQUIRES_OK(ctx, Check(davi_conv2d_wt(P(w), P(&wt), Cout, kh * kw, Cin, StreamOf(ctx)), "davi_conv2d_wt"));
    Tensor* dx;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({B, H, W, Cin}), &dx));
    void* ws; size_t wsb;
    OP_REQUIRES_OK(ctx, ws_.Get(ctx, &ws, &wsb));
    // the input gradient of a stride-1 'same' convolution = the forward kernel on the transposed + flipped weights
    OP_REQUIRES_OK(ctx, Check(davi_conv2d_fwd(P(dy), P(&wt), nullptr, P(dx), B, H, W, Cout, Cin, kh, kw, Cout, Cin, ws,
                                              wsb, StreamOf(ctx)), "davi_conv2d_fwd(input gradient)"));
  }
 private:
  ConvWorkspace ws_;
};
REGISTER_KERNEL_BUILDER(Name("DaviConv2dInputGrad").Device(tf::DEVICE_GPU), DaviConv2dInputGradOp);

REGISTER_OP("DaviConv2dFilterGrad")
    .Input("x: float").Input("dy: float")
    .Attr("kh: int").Attr("kw: int")
    .Output("dkernel_ohwi: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      c->set_output(0, c->UnknownShape());
      return tf::Status::OK();
    });

class DaviConv2dFilterGradOp : public OpKernel {
 public:
  explicit DaviConv2dFilterGradOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("kh", &kh_));
    OP_REQUIRES_OK(c, c->GetAttr("kw", &kw_));
  }
  void Compute(OpKernelContext* ctx) override {
    const Tensor& x = ctx->input(0);
    const Tensor& dy = ctx->input(1);
    const int B = x.dim_size(0), H = x.dim_size(1), W = x.dim_size(2), Cin = x.dim_size(3), Cout = dy.dim_size(3);
    Tensor* dw;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({Cout, kh_, kw_, Cin}), &dw));
    void* ws; size_t wsb;
    OP_REQUIRES_OK(ctx, ws_.Get(ctx, &ws, &wsb));
    OP_REQUIRES_OK(ctx, Check(davi_conv2d_wgrad(P(x), P(dy), P(dw), B, H, W, Cin, Cout, kh_, kw_, Cin, Cout, ws, wsb,
                                                StreamOf(ctx)), "davi_conv2d_wgrad"));
  }
 private:
  int kh_, kw_;
  ConvWorkspace ws_;
};
REGISTER_KERNEL_BUILDER(Name("DaviConv2dFilterGrad").Device(tf::DEVICE_GPU), DaviConv2dFilterGradOp);
