// Minimal stand-in for tensorflow/core/framework/op.h (see op_kernel.h in this directory): REGISTER_OP's builder.
#pragma once
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"

namespace tensorflow {

struct OpDefBuilderStub {
  explicit OpDefBuilderStub(const char*) {}
  OpDefBuilderStub& Input(const char*) { return *this; }
  OpDefBuilderStub& Output(const char*) { return *this; }
  OpDefBuilderStub& Attr(const char*) { return *this; }
  OpDefBuilderStub& Doc(const char*) { return *this; }
  OpDefBuilderStub& SetShapeFn(Status (*fn)(shape_inference::InferenceContext*)) { (void)fn; return *this; }
};

}  // namespace tensorflow

#define REGISTER_OP(name) \
  static ::tensorflow::OpDefBuilderStub DAVI_TFSTUB_CAT(davi_tfstub_op_, __COUNTER__) = ::tensorflow::OpDefBuilderStub(name)
