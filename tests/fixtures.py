"""The reference's own test problems, restated as plain arrays.

  logistic_l1()  — test/test_logistic_l1.jl:12-46  (8×5 data, golden x_star :29)
  sharing()      — test/test_sharing.jl:9-28       (3 blocks of dim 2, golden sum_star :28)
  planted_lasso()— test/test_lasso.jl:15-60        (N,n = 6,3; optimum planted by construction)

The Lasso instance is rebuilt with numpy's RNG: Julia's ``Random.seed!(0)``
stream is not reproducible without Julia, and the reference tests are
RNG-agnostic because x_star is optimal by construction.
"""
import numpy as np

X_STAR_LOGISTIC = np.array([0.0, 0.924160995722576, -1.1343956493097298, 0.0, 0.0])
SUM_STAR_SHARING = np.array([-5.136781609195401, -0.9333333333333327])


def logistic_l1():
    x_class1 = np.array([[5.1, 3.5, 1.4, 0.2, 1.0],
                         [4.9, 3.0, 1.4, 0.2, 1.0],
                         [4.7, 3.2, 1.3, 0.2, 1.0],
                         [4.6, 3.1, 1.5, 0.2, 1.0]])
    x_class2 = np.array([[5.7, 3.0, 4.2, 1.2, 1.0],
                         [5.7, 2.9, 4.2, 1.3, 1.0],
                         [6.2, 2.9, 4.3, 1.3, 1.0],
                         [5.1, 2.5, 3.0, 1.1, 1.0]])
    xs = np.vstack([x_class1, x_class2])
    ys = np.concatenate([np.full(4, 1.0), np.full(4, -1.0)])
    N, n = xs.shape
    L = 0.25 * np.sum(xs * xs, axis=1)          # :39  0.25*norm(xs[i,:])^2
    return dict(A=xs, y=ys, mu=np.ones(N), L=L, lam=1.0 / N, x0=np.ones(n), N=N, n=n,
                x_star=X_STAR_LOGISTIC)


def sharing():
    n, N = 2, 3
    eta = N * 10.0
    dq = np.array([[1.0, 2.0], [-1.0, 3.0], [0.0, 10.0]])
    qlin = np.ones((N, n))
    # :23  L_i = opnorm(Q[i]) + η with Q[i] a LINEAR index into the last Q built
    # (a scalar): Q = diagm(d_i) column-major → Q[1]=d_i[1], Q[2]=0, Q[3]=0.
    # The loop indexes the current Q with i, so L = [|1|, 0, 0] + η = [31, 30, 30].
    L = np.array([abs(dq[0, 0]), 0.0, 0.0]) + eta
    return dict(Qdiag=dq, qlin=qlin, box=(-2.0, 2.0), eta=eta, L=L, N=N, n=n,
                g_hi=np.ones(n), x0=np.zeros(n), sum_star=SUM_STAR_SHARING)


def planted_lasso(seed=0, N=6, n=3, p=2, rho=10.0, lam=1.0):
    rs = np.random.RandomState(seed)
    y_star = rs.rand(N)
    y_star /= np.linalg.norm(y_star)
    Cm = rs.rand(N, n) * 2 - 1
    CTy = np.abs(Cm.T @ y_star)
    perm = np.argsort(-CTy, kind="stable")
    alpha = np.zeros(n)
    for i in range(n):
        if i < p:
            alpha[perm[i]] = lam / CTy[perm[i]]
        else:
            alpha[perm[i]] = lam if CTy[perm[i]] < 0.1 * lam else lam * rs.rand() / CTy[perm[i]]
    A = Cm * alpha[None, :]
    x_star = np.zeros(n)
    for i in range(p):
        x_star[perm[i]] = rs.rand() * rho / np.sqrt(p) * np.sign(A[:, perm[i]] @ y_star)
    b = A @ x_star + y_star

    def cost(x):
        return np.linalg.norm(A @ x - b) ** 2 / 2 + lam * np.abs(x).sum()

    L = N * np.sum(A * A, axis=1)               # :55  opnorm(tempA)^2 * N
    return dict(A=A, b=b, scale=np.full(N, float(N)), L=L, lam=lam, x0=np.zeros(n), N=N, n=n,
                x_star=x_star, f_star=cost(x_star), cost=cost)
