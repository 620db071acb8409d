"""A SECOND, independent restatement of the reference's loops — numpy array statements written from the Julia
sources, one statement per Julia broadcast statement — used only to cross-check the C oracle (oracle/ciao_oracle.c)
step by step (tests/test_oracle_cross.py).  The reference holds no per-step golden trajectories and Julia is not
available here, so two independent transcriptions agreeing to rounding is the strongest per-step pin this image allows.
Test infrastructure only; nothing in the product imports it.

Operators follow ProximalOperators.jl 0.14; rows are 1×d (test_lasso.jl:53-54, test_logistic_l1.jl:36),
sharing blocks are Sum(Quadratic(diagm(q_i), c_i), SqrDistL2(IndBox(lo, hi), η)) (test_sharing.jl:18-22).
"""
import numpy as np


# ---- f_i -------------------------------------------------------------------------------------------------
class LeastSquaresRow:
    """LeastSquares(A[i:i, :], b[i:i], λ):  f = (λ/2)(a·x − b)²;  gradient!: res = A x − b; y = Aᴴ res; y .*= λ."""

    def __init__(self, a, b, lam):
        self.a, self.b, self.lam = np.asarray(a, float), float(b), float(lam)

    def gradient(self, x):
        res = self.a @ x - self.b
        y = self.a * res
        y *= self.lam
        return y, (self.lam / 2) * res * res


class LogisticRow:
    """Precompose(LogisticLoss([y], μ), row, 1.0):  f = μ log(1 + exp(−y·(a·x)));  ∇ = aᵀ(−μ y / (1 + exp(y·(a·x))))."""

    def __init__(self, a, y, mu):
        self.a, self.y, self.mu = np.asarray(a, float), float(y), float(mu)

    def gradient(self, x):
        u = self.a @ x
        with np.errstate(over="ignore", divide="ignore"):        # exp overflows to Inf / 1/0 = Inf exactly as in Julia
            e = np.exp(self.y * u)
            return self.a * (-self.mu * self.y / (1 + e)), self.mu * np.log(1 + 1 / e)


class DiagQuadPlusSqrDist:
    """Sum(Quadratic(diagm(q), c), SqrDistL2(IndBox(lo, hi), η)):  ∇ = (q ∘ x + c) + η (x − clamp(x, lo, hi))."""

    def __init__(self, q, c, lo, hi, eta):
        self.q, self.c, self.lo, self.hi, self.eta = np.asarray(q, float), np.asarray(c, float), lo, hi, float(eta)

    def gradient(self, x):
        g1 = self.q * x + self.c
        p = np.clip(x, self.lo, self.hi)
        g2 = self.eta * (x - p)
        f = 0.5 * np.dot(x, self.q * x) + np.dot(self.c, x) + (self.eta / 2) * np.dot(x - p, x - p)
        return g1 + g2, f


# ---- g ---------------------------------------------------------------------------------------------------
class NormL1:
    def __init__(self, lam):
        self.lam = float(lam)

    def prox(self, x, gamma):
        gl = gamma * self.lam                         # y = x + (x ≤ −gl ? gl : (x ≥ gl ? −gl : −x))
        return x + np.where(x <= -gl, gl, np.where(x >= gl, -gl, -x))


class IndBox:
    def __init__(self, lo, hi):
        self.lo, self.hi = lo, hi

    def prox(self, x, gamma):
        return np.clip(x, self.lo, self.hi)


class Zero:
    def prox(self, x, gamma):
        return x.copy()


def static_batches(N, r):
    """Finito_basic.jl:52-57 (1-based row numbers)."""
    d = N // r
    ind = [list(range(r * i + 1, r * (i + 1) + 1)) for i in range(d)]
    if r * d < N:
        ind.append(list(range(r * d + 1, N + 1)))
    return ind


# ---- SVRG_basic.jl ---------------------------------------------------------------------------------------
class SVRG:
    def __init__(self, F, g, x0, gamma, plus=False):          # :57-68
        self.F, self.g, self.N, self.gamma, self.plus = F, g, len(F), gamma, plus
        self.av = np.zeros_like(x0, dtype=float)
        for f in F:
            grad, _ = f.gradient(x0)
            grad /= self.N
            self.av += grad
        self.z_full = np.array(x0, float)
        self.z = np.zeros_like(self.z_full)
        self.w = np.array(x0, float)

    def epoch(self, idx1):                                       # :71-96
        m = len(idx1)
        for i in idx1:
            temp, _ = self.F[i - 1].gradient(self.z_full)
            gtemp, _ = self.F[i - 1].gradient(self.w)
            temp -= gtemp
            temp -= self.av
            temp *= self.gamma
            temp += self.w
            self.w = self.g.prox(temp, self.gamma)
            self.z += self.w
        self.z_full = self.z / m
        if not self.plus:
            self.w = self.z_full.copy()
        self.z = np.zeros_like(self.z)
        self.av = self.z.copy()
        for f in self.F:
            grad, _ = f.gradient(self.z_full)
            grad /= self.N
            self.av += grad


# ---- SAGA_basic.jl ---------------------------------------------------------------------------------------
class SAGA:
    def __init__(self, F, g, x0, gamma, sag=False):              # :41-50
        self.F, self.g, self.N, self.gamma, self.sag = F, g, len(F), gamma, sag
        self.s = [f.gradient(x0)[0] for f in F]
        tot = self.s[0].copy()
        for v in self.s[1:]:
            tot = tot + v
        self.av = tot / self.N
        self.z = g.prox((1 - gamma) * np.asarray(x0, float), gamma)

    def step(self, i1):                                          # :53-68
        i, N = i1 - 1, self.N
        grad, _ = self.F[i].gradient(self.z)
        if self.sag:
            self.av = self.av + (grad - self.s[i]) / N
            w = self.z - self.gamma * self.av
        else:
            w = self.z - self.gamma * (grad - self.s[i] + self.av)
            self.av = self.av + (grad - self.s[i]) / N
        self.z = self.g.prox(w, self.gamma)
        self.s[i] = grad.copy()


# ---- Finito_basic.jl -------------------------------------------------------------------------------------
class Finito:
    def __init__(self, F, g, x0, gam):                           # :76-87
        self.F, self.g, self.N, self.gam = F, g, len(F), np.asarray(gam, float)
        N = self.N
        x0 = np.asarray(x0, float)
        self.s = [x0 - self.gam[i] / N * F[i].gradient(x0)[0] for i in range(N)]
        self.hat = 1 / np.sum(1 / self.gam)
        tot = self.s[0] / self.gam[0]
        for i in range(1, N):
            tot = tot + self.s[i] / self.gam[i]
        self.av = self.hat * tot
        self.z = g.prox(self.av, self.hat)

    def step(self, batch1):                                      # :110-119
        for i1 in batch1:
            i = i1 - 1
            t, _ = self.F[i].gradient(self.z)
            t *= -(self.gam[i] / self.N)
            t += self.z
            self.av = self.av + (t - self.s[i]) * (self.hat / self.gam[i])
            self.s[i] = t.copy()
        self.z = self.g.prox(self.av, self.hat)


# ---- Finito_LFinito.jl -----------------------------------------------------------------------------------
class LFinito:
    def __init__(self, F, g, x0, gam, batch=1):                  # :40-76
        self.F, self.g, self.N, self.gam = F, g, len(F), np.asarray(gam, float)
        self.hat = 1 / np.sum(1 / self.gam)
        self.ind = static_batches(self.N, batch)
        self.av = np.array(x0, float)
        for f in F:
            grad, _ = f.gradient(np.asarray(x0, float))
            grad *= self.hat / self.N
            self.av -= grad
        self.z = np.zeros_like(self.av)
        self.z_full = np.zeros_like(self.av)

    def outer(self, order1):                                     # :78-103
        N, hat = self.N, self.hat
        self.z_full = self.g.prox(self.av, hat)
        self.av = self.z_full.copy()
        for f in self.F:
            grad, _ = f.gradient(self.z_full)
            self.av -= (hat / N) * grad
        for j in order1:
            self.z = self.g.prox(self.av, hat)
            for i1 in self.ind[j - 1]:
                i = i1 - 1
                grad, _ = self.F[i].gradient(self.z_full)
                self.av += (hat / N) * grad
                grad, _ = self.F[i].gradient(self.z)
                self.av -= (hat / N) * grad
                self.av += (hat / self.gam[i]) * (self.z - self.z_full)


# ---- ProShI_basic.jl -------------------------------------------------------------------------------------
class ProShI:
    def __init__(self, F, g, x0, gam):                           # :76-90
        self.F, self.g, self.N, self.gam = F, g, len(F), np.asarray(gam, float)
        N = self.N
        x0 = np.asarray(x0, float)
        self.s = [x0 - self.gam[i] / N * F[i].gradient(x0)[0] for i in range(N)]
        self.hat = float(np.sum(self.gam))
        tot = self.s[0].copy()
        for v in self.s[1:]:
            tot = tot + v
        self.av = tot
        self.z = g.prox(self.av, self.hat)
        self.z -= self.av
        self.z /= self.hat

    def step(self, batch1):                                      # :110-124
        for i1 in batch1:
            i = i1 - 1
            self.av -= self.s[i]
            self.s[i] = self.s[i] + self.gam[i] * self.z
            t, _ = self.F[i].gradient(self.s[i])
            t *= -(self.gam[i] / self.N)
            t += self.s[i]
            self.av += t
            self.s[i] = t.copy()
        self.z = self.g.prox(self.av, self.hat)
        self.z -= self.av
        self.z /= self.hat

    def solution(self):                                          # :127-132 (mutates the table)
        for i in range(self.N):
            self.s[i] = self.s[i] + self.gam[i] * self.z
        return self.s


# ---- Finito_adaptive.jl ----------------------------------------------------------------------------------
class FinitoAdaptive:
    def __init__(self, F, g, x0, alpha=0.999, tol_b=1e-9, perturb=None):       # :59-99; perturb(i, t) = rand(t*[-1,1], size(x0))
        self.F, self.g, self.N, self.alpha, self.tol_b = F, g, len(F), alpha, tol_b
        N = self.N
        x0 = np.asarray(x0, float)
        self.s, self.gf, self.fi_x = [], [], np.zeros(N)
        for i in range(N):
            grad, fx = F[i].gradient(x0)
            self.gf.append(grad)
            self.fi_x[i] = fx
            self.s.append(x0.copy())
        self.gam = np.zeros(N)
        for i in range(N):
            xeps = x0 + 1.0
            grad_eps, _ = F[i].gradient(xeps)
            nmg = np.linalg.norm(grad_eps - self.gf[i])
            t = 1
            while nmg < np.finfo(float).eps:                      # :77-83, the draw is the caller's
                assert perturb is not None, "∇f_i(x0 + 1) == ∇f_i(x0) and no perturbation source"
                xeps = x0 + perturb(i + 1, t)
                grad_eps, _ = F[i].gradient(xeps)
                nmg = np.linalg.norm(grad_eps - self.gf[i])
                t *= 2
            L_int = nmg / (t * np.sqrt(len(x0)))
            L_int /= N
            self.gam[i] = alpha / L_int
        self.hat = 1 / np.sum(1 / self.gam)
        tot_s = self.s[0] / self.gam[0]
        tot_g = self.gf[0].copy()
        for i in range(1, N):
            tot_s = tot_s + self.s[i] / self.gam[i]
            tot_g = tot_g + self.gf[i]
        self.av = self.hat * (tot_s - tot_g / N)
        self.z = g.prox(self.av, self.hat)
        self.backtracks = 0

    def step(self, i1):                                          # :101-160; False ⇔ `return nothing`
        i, N = i1 - 1, self.N
        res = self.z - self.s[i]
        while True:
            if self.gam[i] < self.tol_b / N:
                return False
            _, fi_z = self.F[i].gradient(self.z)
            fi_model = self.fi_x[i] + np.dot(self.gf[i], res) + (0.5 * N * self.alpha / self.gam[i]) * (np.linalg.norm(res) ** 2)
            tol = 10 * np.finfo(float).eps * (1 + abs(fi_z))
            if fi_z <= fi_model + tol:
                break
            gam_b = self.gam[i]
            self.gam[i] *= 0.8
            self.av = self.av / self.hat
            self.av = self.av + self.s[i] / self.gam[i]
            self.av = self.av - self.s[i] / gam_b
            self.hat = 1 / (1 / self.hat + 1 / self.gam[i] - 1 / gam_b)
            self.av = self.av * self.hat
            self.z = self.g.prox(self.av, self.hat)
            res = self.z - self.s[i]
            self.backtracks += 1
        self.av = self.av + (self.hat / self.gam[i]) * (self.z - self.s[i])
        self.s[i] = self.z.copy()
        self.av = self.av + (self.hat / N) * self.gf[i]
        self.gf[i], self.fi_x[i] = self.F[i].gradient(self.z)
        self.av = self.av - (self.hat / N) * self.gf[i]
        self.z = self.g.prox(self.av, self.hat)
        return True
