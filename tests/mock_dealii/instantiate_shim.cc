// instantiates every template of contrib/dealii/b200mf_dealii.h against the mock (tests/test_shim_syntax.py)
#include <b200mf_dealii.h>

struct MaterialMock
{
  double mu = 1, nu = 0.3, rho = 1, rho_alpha = 1;
  int    formulation = 1;
};

template class B200MF::NeoHookOperator<3, double>;
template class B200MF::NeoHookOperator<3, float>;
template class B200MF::LaplaceOperator<3, double>;
template class B200MF::LaplaceOperator<3, float>;
template b200mf_ctx *B200MF::make_context<3, 3, double>(const dealii::MatrixFree<3, double> &, b200mf_comm *);
template b200mf_ctx *B200MF::make_context<3, 1, float>(const dealii::MatrixFree<3, float> &, b200mf_comm *);

void use()
{
  B200MF::NeoHookOperator<3, double> op;
  auto m = std::make_shared<MaterialMock>();
  op.set_material(m, m);
  dealii::LinearAlgebra::distributed::Vector<double> u, v;
  std::shared_ptr<const dealii::MatrixFree<3, double>> mf;
  op.initialize(mf, mf, u, "tensor2", B200MF::make_communicator(0));
  op.cache();
  op.vmult(v, u);
  op.compute_diagonal();
  op.precondition_Jacobi(v, u, 0.65);
  (void)op.el(0, 0);
  B200MF::DeviceVector d(nullptr), s(nullptr);
  std::vector<double> scratch;
  s.upload(u, scratch);
  op.cache(s);
  op.vmult(d, s);
  d.download(v, scratch);
  d.sadd(1.0, 1.0, s);
  (void)d.l2_norm();
  B200MF::LaplaceOperator<3, float> lap;
  std::shared_ptr<const dealii::MatrixFree<3, float>> mff;
  lap.initialize(mff, 0.1);
  dealii::LinearAlgebra::distributed::Vector<float> a, b;
  lap.vmult(a, b);
  lap.vmult_interface_down(a, b);
  lap.compute_diagonal();
  lap.set_delta_t(0.2);
  (void)lap.m();
  std::shared_ptr<const dealii::MatrixFree<3, double>> mfd;
  B200MF::BoundaryProjector<3> projector(7, mfd);
  projector.project(u);
  projector.project(d);
  (void)projector.last_step();
}
