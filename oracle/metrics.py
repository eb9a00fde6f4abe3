"""Evaluation metrics — oracle restatement of main_diffusion_linear.py:84-117 and
main_diffusion_scatterometry.py:72-101 (histogram KL of posterior samples).  TEST INFRASTRUCTURE ONLY.
"""
import numpy as np
import scipy.special


def hist_sum(sample_sets, bins, ranges):
    """sum over repeats of np.histogramdd(x, bins=bins, range=ranges)[0]   (:85-91)"""
    total = None
    for x in sample_sets:
        h, _ = np.histogramdd(np.asarray(x), bins=bins, range=ranges)
        total = h if total is None else total + h
    return total


def kl2(hist_true_sum, hist_model_sum, epsilon=1e-10):
    """normalise, add epsilon, re-normalise, sum(rel_entr(true, model))   (:108-116)"""
    p = hist_true_sum / hist_true_sum.sum()
    q = hist_model_sum / hist_model_sum.sum()
    p = p + epsilon
    q = q + epsilon
    p /= p.sum()
    q /= q.sum()
    return float(np.sum(scipy.special.rel_entr(p, q)))
