"""Random-walk Metropolis on the scatterometry posterior — oracle restatement of anneal_to_energy
(models/SNF.py:250-275, langevin_prop=False) with E = get_log_posterior (utils_scatterometry.py:30-38), as
generate_scatterometry_ground_truth.py:26-29 uses it.  TEST INFRASTRUCTURE ONLY.
"""
import torch

from . import scatterometry as scat


def anneal_to_energy(params, x0, y, noise, unif, noise_std, a=scat.A_NOISE, b=scat.B_NOISE, lambd_bd=scat.LAMBD_BD):
    """x0 (n,3) start points; y (n,23) the observation of each chain; noise (S,n,3) standard normals and unif (S,n)
    uniforms in the order the reference draws them (randn_like :259, rand_like :266).
    Returns (x_final, E(x_final) - E(x0))."""
    x = x0.clone()
    e0 = scat.energy(params, x, y, a, b, lambd_bd)
    e = e0.clone()
    for z, u in zip(noise, unif):
        x_prop = x + noise_std * z
        e_prop = scat.energy(params, x_prop, y, a, b, lambd_bd)
        acc = u < torch.exp(-e_prop + e)          # e = E(x_curr): the reference re-evaluates it every step
        x = torch.where(acc[:, None], x_prop, x)
        e = torch.where(acc, e_prop, e)
    return x, e - e0
