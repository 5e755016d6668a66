"""SDEs written the way one writes them for mattja/sdeint: plain Python callables f(y, t), G(y, t).  Shared by
oracle/make_golden.py -- which hands them to the UNMODIFIED reference -- and by the GPU tests, which hand the same
objects to sdeint_b200 (where they are traced into device code): the two must agree to 1e-12."""
import numpy as np


def heston():
    """Full-truncation Heston: y = (S, v), correlated noise through a Cholesky factor."""
    r, kappa, theta, xi, rho = 0.03, 2.0, 0.04, 0.3, -0.7

    def f(y, t):
        return np.array([r * y[0], kappa * (theta - np.maximum(y[1], 0.0))])

    def G(y, t):
        sv = np.sqrt(np.maximum(y[1], 0.0))
        return np.array([[sv * y[0], 0.0], [xi * sv * rho, xi * sv * np.sqrt(1.0 - rho ** 2)]])
    return f, G, np.array([1.0, 0.04])


def piecewise():
    """Thresholds and masks: np.where, np.sign, comparisons, a step in time."""
    def f(y, t):
        return np.array([np.where(y > 0.3, -2.0 * y, 0.5 - y)[0] + 0.2 * np.heaviside(t - 0.25, 0.5),
                         -np.sign(y[1]) * np.sqrt(np.maximum(abs(y[1]), 1e-3)) + 0.1 * y[0]])

    def G(y, t):
        return np.array([[0.3 * (y[0] > 0) * y[0], 0.1], [0.05 * (y[1] <= 0.1), 0.2 + 0.1 * np.tanh(y[0])]])
    return f, G, np.array([0.5, -0.2])


def oscillator():
    """A noisy nonlinear oscillator with the preallocated-result idiom and time-dependent forcing."""
    def f(y, t):
        dy = np.zeros(3)
        dy[0] = y[1]
        dy[1] = -y[0] - 0.3 * y[1] - 0.5 * y[0] ** 3 + 0.4 * np.cos(1.7 * t) * y[2]
        dy[2] = -0.2 * y[2] + np.arctan(y[0])
        return dy

    def G(y, t):
        B = np.zeros((3, 2))
        B[1, 0] = 0.25 * (1.0 + 0.5 * np.sin(y[0]))
        B[1, 1] = 0.05 * y[2]
        B[2, 1] = 0.15 * np.exp(-0.5 * t)
        return B
    return f, G, np.array([0.8, 0.0, 0.1])


def linear_np_dot(d=5, m=4, seed=11):
    """Roessler (7.4)-type system with np.dot; also the reference's list-of-columns form of G."""
    rng = np.random.default_rng(seed)
    A = -np.eye(d) + 0.1 * rng.standard_normal((d, d))
    B = 0.2 * rng.standard_normal((m, d, d))

    def f(y, t):
        return np.dot(A, y)

    def G(y, t):
        return np.dot(B, y).T
    cols = [(lambda y, t, k=k: np.dot(B[k], y)) for k in range(m)]
    return f, G, np.ones(d), cols


SYSTEMS = {"heston": heston, "piecewise": piecewise, "oscillator": oscillator, "linear_np_dot": linear_np_dot}
