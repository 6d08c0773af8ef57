"""solver_GP facade with the reference's interface and print strings (src/solver.py:41-206).

Differences, all deliberate: (i) no matplotlib dependency -- the plotting methods import it lazily and
raise a clear error when it is absent; (ii) the upstream bug in get_sample / get_sample_IP (self passed
twice, src/solver.py:86,110) is fixed; (iii) error metrics use numpy."""
import numpy as onp

from .PDEs import Nonlinear_elliptic2d, Burgers, Eikonal
from .InverseProblems import Darcy_flow2d


def _plt():
    try:
        import matplotlib.pyplot as plt
        return plt
    except Exception as e:  # pragma: no cover
        raise RuntimeError("plotting needs matplotlib, which is not installed in this environment") from e


class solver_GP(object):
    def __init__(self, cfg=None, PDE_type="Nonlinear_elliptic"):
        self.config = cfg
        self.PDE_type = PDE_type

    def set_equation(self, bdy=None, rhs=None, domain=onp.array([[0, 1], [0, 1]]), print_option=True):
        if self.PDE_type == "Nonlinear_elliptic":
            self.eqn = Nonlinear_elliptic2d(alpha=self.config.alpha, m=self.config.m, bdy=bdy, rhs=rhs, domain=domain)
            if print_option:
                print('\n Solver started')
                print('[Equation type] Nonlinear elliptic equation')
                print('[Equation form] - \\Delta u + alpha*u^m = f')
                print(f'[Equation domain] [{domain[0,0]},{domain[0,1]}]*[{domain[1,0]},{domain[1,1]}]')
                print(f'[Equation parameter] alpha = {self.config.alpha}, m = {self.config.m}')
                print('[Equation data] Right hand side and boundary values set by the user')
        elif self.PDE_type == "Burgers":
            self.eqn = Burgers(alpha=self.config.alpha, nu=self.config.nu, bdy=bdy, rhs=rhs, domain=domain)
            if print_option:
                print('\n Solver started')
                print('[Equation type] Burgers equation')
                print('[Equation form] u_t+ alpha u u_x- nu u_xx=0')
                print(f'[Equation domain] [{domain[0,0]},{domain[0,1]}]*[{domain[1,0]},{domain[1,1]}]')
                print(f'[Equation parameter] alpha = {self.config.alpha}, m = {self.config.nu}')
                print('[Equation data] Right hand side and boundary values set by the user')
        elif self.PDE_type == "Eikonal":
            self.eqn = Eikonal(eps=self.config.eps, bdy=bdy, rhs=rhs, domain=domain)
            if print_option:
                print('\n Solver started')
                print('[Equation type] Eikonal equation')
                print('[Equation form] |grad u|^2 = f + eps*Delta u')
                print(f'[Equation domain] [{domain[0,0]},{domain[0,1]}]*[{domain[1,0]},{domain[1,1]}]')
                print(f'[Equation parameter] eps = {self.config.eps}')
                print('[Equation data] Right hand side and boundary values set by the user')
        elif self.PDE_type == "Darcy_flow2d":
            self.eqn = Darcy_flow2d(bdy=bdy, rhs=rhs, domain=domain)
            if print_option:
                print('\n Solver started')
                print('[Inverse problem type] Darcy flow 2d')
                print('[Inverse problem form] -div(a grad u) = f, infer a from f and some observed u')
                print(f'[Equation domain] [{domain[0,0]},{domain[0,1]}]*[{domain[1,0]},{domain[1,1]}]')
                print('[Equation data] Right hand side and boundary values set by the user')
        else:
            raise ValueError(f"unknown PDE_type {self.PDE_type!r}")

    def get_sample(self, X_domain, X_boundary, print_option=True):
        self.eqn.get_sampled_points(X_domain, X_boundary)
        if print_option:
            print('[Sample points] Collocation points sampled, specified by the user')
            print(f'[Sample points] N_domain = {self.eqn.N_domain}, N_boundary = {self.eqn.N_boundary}')

    def auto_sample(self, N_domain, N_boundary, sampled_type='random', print_option=True):
        self.eqn.sampled_pts(N_domain, N_boundary, sampled_type=sampled_type)
        if print_option:
            print(f'[Sample points] Collocation points sampled, type {sampled_type}')
            print(f'[Sample points] N_domain = {self.eqn.N_domain}, N_boundary = {self.eqn.N_boundary}')

    def show_sample(self):
        plt = _plt()
        fig = plt.figure()
        ax = fig.add_subplot(111)
        ax.scatter(self.eqn.X_domain[:, 0], self.eqn.X_domain[:, 1], marker="x", label='Interior nodes')
        ax.scatter(self.eqn.X_boundary[:, 0], self.eqn.X_boundary[:, 1], marker="x", label='Boundary nodes')
        ax.legend(loc="upper right")
        plt.title('Collocation points')

    def get_sample_IP(self, X_domain, X_boundary, X_data, print_option=True):
        self.eqn.get_sampled_points(X_domain, X_boundary, X_data)
        if print_option:
            print('[Sample points] Collocation points sampled, specified by the user')
            print(f'[Sample points] N_domain = {self.eqn.N_domain}, N_boundary = {self.eqn.N_boundary}, N_data = {self.eqn.N_data}')

    def auto_sample_IP(self, N_domain, N_boundary, N_data, sampled_type='random', print_option=True):
        self.eqn.sampled_pts(N_domain, N_boundary, N_data, sampled_type=sampled_type)
        if print_option:
            print(f'[Sample points] Collocation points sampled, type {sampled_type}')
            print(f'[Sample points] N_domain = {self.eqn.N_domain}, N_boundary = {self.eqn.N_boundary}, N_data = {self.eqn.N_data}')

    def show_sample_IP(self):
        plt = _plt()
        fig = plt.figure()
        ax = fig.add_subplot(111)
        ax.scatter(self.eqn.X_domain[:, 0], self.eqn.X_domain[:, 1], label='Interior nodes')
        ax.scatter(self.eqn.X_boundary[:, 0], self.eqn.X_boundary[:, 1], label='Boundary nodes')
        ax.scatter(self.eqn.X_domain[:self.eqn.N_data, 0], self.eqn.X_domain[:self.eqn.N_data, 1], label='Data nodes')
        ax.legend(loc="upper right")
        plt.title('Collocation and data points')

    def get_observed_data(self, data_u, noise_level, print_option=True):
        self.eqn.get_observation(data_u, noise_level)
        if print_option:
            print('[Observed Data] Get observed data from solving the PDE using FD and interpolation')
            print(f'[Observed Data] Noise level {noise_level}')

    def solve(self, method='elimination', pen_lambda=1e-10, print_option=True):
        cfg = self.config
        if print_option:
            print('[Kernel] ' + cfg.kernel)
            print(f'[Kernel parameter]: {cfg.kernel_parameter}')
        self.eqn.Gram_matrix(kernel=cfg.kernel, kernel_parameter=cfg.kernel_parameter, nugget=cfg.nugget, nugget_type=cfg.nugget_type)
        if print_option:
            print(f'[Gram matrix] Finish assembly of the Gram matrix, nugget {cfg.nugget}, type {cfg.nugget_type}')
        self.eqn.Gram_Cholesky()
        if print_option:
            print('[Gram matrix] Finish Cholesky factorization of the Gram matrix')
            print('[Gauss Newton] Start Gauss Newton iteration')
            print(f'[Gauss Newton] {method} approaches')
        if method == 'elimination':
            self.eqn.GN_method(max_iter=cfg.GNsteps, step_size=cfg.step_size, initial_sol=cfg.initial_sol, print_hist=cfg.print_hist)
        elif method == 'relaxation':
            self.eqn.GN_relaxed_method(max_iter=cfg.GNsteps, step_size=cfg.step_size, initial_sol=cfg.initial_sol, pen_lambda=pen_lambda, print_hist=cfg.print_hist)
        if print_option:
            print('[Gauss Newton] Gauss Newton iteration finished')

    def show_loss_hist(self):
        plt = _plt()
        plt.figure()
        plt.plot(onp.arange(self.eqn.max_iter + 1), self.eqn.loss_hist)
        plt.yscale("log")
        plt.title('Loss function history')
        plt.xlabel('Gauss-Newton step')

    def collocation_pts_err(self, truth, print_option=True):
        if print_option:
            print('[Calculating collocation errors...]')
        self.pts_err_all = abs(truth - self.eqn.sol_sampled_pts)
        self.pts_max_err = onp.max(self.pts_err_all)
        self.pts_L2_err = onp.sqrt(onp.sum(self.pts_err_all ** 2) / (self.eqn.N_domain))
        if print_option:
            print(f'[Collocation point error] Max error {self.pts_max_err}')
            print(f'[Collocation point error] L2 error {self.pts_L2_err}')

    def test(self, X_test, print_option=True):
        if print_option:
            print(f'[Testing...] Number of test points: {X_test.shape[0]}')
        self.eqn.extend_sol(X_test)

    def get_test_error(self, truth, print_option=True):
        self.truth = truth
        self.test_err_all = abs(truth - self.eqn.extended_sol)
        self.test_max_err = onp.max(self.test_err_all)
        self.test_L2_err = onp.sqrt(onp.sum(self.test_err_all ** 2) / (self.eqn.N_test))
        if print_option:
            print(f'[Test error] Max error {self.test_max_err}')
            print(f'[Test error] L2 error {self.test_L2_err}')

    def contour_of_test_err(self, XX, YY):
        plt = _plt()
        fig = plt.figure()
        ax = fig.add_subplot(111)
        c = ax.contourf(XX, YY, self.test_err_all.reshape(XX.shape), 50, cmap=plt.cm.coolwarm)
        self.XX, self.YY = XX, YY
        plt.xlabel('$x_1$')
        plt.ylabel('$x_2$')
        plt.title('Contour of errors')
        fig.colorbar(c)
        plt.show()
