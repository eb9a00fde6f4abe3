"""The linear-Gaussian inverse problem of the reference (`linear_problem.py:5-65`): y = A x + b + N(0, scale I),
x ~ N(0, I), with its closed-form posterior and posterior score — the `forward_model` argument of the training and
evaluation loops and the PINNLoss initial condition (`main_diffusion_linear.py:151-156`).  Same attributes and methods;
every method follows the device of its argument, so it works on the CUDA tensors the fused paths hand it.
"""
import torch
from torch.distributions import MultivariateNormal

device = 'cuda' if torch.cuda.is_available() else 'cpu'


class LinearForwardProblem:
    def __init__(self):
        self.epsilon = 1e-6
        self.xdim = 2
        self.ydim = 2
        self.A = torch.tensor([[1.0, 0.5], [0.0, 1.0]])     # shear by 0.5 along x
        self.b = torch.tensor([0.3, 0.5])                   # translation
        self.scale = 0.3
        eye_x, eye_y = torch.eye(self.xdim), torch.eye(self.ydim)
        self.Sigma = self.scale * eye_y
        self.Lam = eye_x
        self.Sigma_inv = eye_y / self.scale
        self.Sigma_y_inv = torch.linalg.inv(self.Sigma + self.A @ self.Lam @ self.A.T + self.epsilon * eye_y)
        self.mu = torch.zeros(self.xdim)

    def __call__(self, *args, **kwargs):
        return self.forward(args[0])

    def forward(self, x):
        return x @ self.A.to(x).T + self.b.to(x)

    def get_likelihood(self, x):
        return MultivariateNormal(self.A.to(x) @ x + self.b.to(x), self.Sigma.to(x))

    def get_evidence(self):
        return MultivariateNormal(self.A @ self.mu + self.b, self.Sigma + self.A @ self.Lam @ self.A.T)

    def _gain(self):
        return self.Lam @ self.A.T @ self.Sigma_y_inv                     # (xdim, ydim)

    def posterior_mean(self, y):
        """mean of p(x | y) for y (ydim,) or a batch (n, ydim)"""
        y_res = y - (self.A @ self.mu + self.b).to(y)
        return y_res @ self._gain().to(y).T

    def posterior_cov(self):
        return self.Lam - self._gain() @ self.A @ self.Lam

    def get_posterior(self, y, device=device):
        y = torch.as_tensor(y, dtype=torch.float32).cpu()
        return MultivariateNormal(self.posterior_mean(y).to(device), self.posterior_cov().to(device))

    def log_posterior(self, xs, ys, epsilon=1e-6):
        """1/2 (x - m)^T C^-1 (x - m), (n, 1) — the value the reference returns under this name (linear_problem.py:48-58),
        reproduced expression by expression: its batch mean is  y_res @ (A^T Sigma_y^-1)  (NOT the posterior mean of
        `get_posterior`, which is y_res @ (Sigma_y^-1 A): an upstream quirk; the only caller is the SNF baseline,
        main_baselines_linear.py:210) and its covariance is  Lam - A^T Sigma_y^-1 A.  Golden: tests/golden/linear_log_posterior.npz."""
        y_res = ys - (self.A @ self.mu + self.b).to(ys)
        mean = y_res @ (self.A.T @ self.Sigma_y_inv).to(ys)
        x_res = xs - mean
        cov = (self.Lam - self.A.T @ self.Sigma_y_inv @ self.A).to(xs)
        cov_inv = torch.linalg.inv(cov + epsilon * torch.eye(self.xdim).to(xs))
        return (0.5 * (x_res @ cov_inv) * x_res).sum(1, keepdim=True)

    def score_posterior(self, x, y):
        """grad_x log p(x | y) = -x + A^T Sigma^-1 (y - A x - b)"""
        A = self.A.to(x)
        y_res = y - (x @ A.T + self.b.to(x))
        return -x + (y_res @ self.Sigma_inv.to(x).T) @ A
