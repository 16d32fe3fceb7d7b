// Minimal stand-in for "ceres/autodiff_cost_function.h" (TEST INFRASTRUCTURE): only what PoseGraph3dErrorTerm::Create
// (REF/include/PoseGraph3dError.h:56-61) needs to compile.  The functor itself is evaluated by oracle/ref_functor.cpp.
#ifndef REF_SHIM_CERES_AUTODIFF_
#define REF_SHIM_CERES_AUTODIFF_
namespace ceres {
class CostFunction { public: virtual ~CostFunction() {} };
template <typename Functor, int kNumResiduals, int N0, int N1, int N2, int N3>
class AutoDiffCostFunction : public CostFunction {
 public:
  explicit AutoDiffCostFunction(Functor* f) : f_(f) {}
  ~AutoDiffCostFunction() { delete f_; }
  const Functor& functor() const { return *f_; }
 private:
  Functor* f_;
};
}  // namespace ceres
#endif
