// Minimal stand-in for tensorflow/core/framework/shape_inference.h (see op_kernel.h in this directory).
#pragma once
#include <initializer_list>
#include "tensorflow/core/framework/op_kernel.h"

namespace tensorflow {
namespace shape_inference {

struct DimensionHandle {
  DimensionHandle() = default;
  DimensionHandle(int64) {}
};
struct ShapeHandle {};
struct DimensionOrConstant {
  DimensionOrConstant(DimensionHandle) {}
  DimensionOrConstant(int64) {}
  DimensionOrConstant(int) {}
};

class InferenceContext {
 public:
  ShapeHandle input(int idx) const { (void)idx; return ShapeHandle(); }
  void set_output(int idx, ShapeHandle shape) { (void)idx; (void)shape; }
  Status WithRank(ShapeHandle shape, int64 rank, ShapeHandle* out) { (void)shape; (void)rank; (void)out; return Status::OK(); }
  DimensionHandle Dim(ShapeHandle s, int64 idx) { (void)s; (void)idx; return DimensionHandle(); }
  ShapeHandle MakeShape(std::initializer_list<DimensionOrConstant> dims) { (void)dims; return ShapeHandle(); }
  ShapeHandle Vector(DimensionOrConstant dim) { (void)dim; return ShapeHandle(); }
  ShapeHandle Scalar() { return ShapeHandle(); }
  ShapeHandle UnknownShape() { return ShapeHandle(); }
  ShapeHandle UnknownShapeOfRank(int64 rank) { (void)rank; return ShapeHandle(); }
};

}  // namespace shape_inference
}  // namespace tensorflow
