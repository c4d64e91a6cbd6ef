"""Test-only numpy stepper with the calling convention of the product's slab stepper: one RK4 step of
the rows [row0, row1) of a slab, reading the whole local array.  Elementwise numpy arithmetic, so a
node's value is independent of how the grid is cut -- the property the partition tests check."""

import numpy as np

WEIGHTS = {3: ([1, -4, 1], 1.0), 5: ([-1, 16, -60, 16, -1], 12.0), 7: ([2, -27, 270, -980, 270, -27, 2], 180.0)}


def numpy_stepper(plan, cols, dx, dt, order, coeffs):
    num, scale = WEIGHTS[order]
    k = (order - 1) // 2
    w = np.array(num, dtype=float) / (scale * (dx * dx))
    c = np.asarray(coeffs, dtype=float)

    def lap(y):
        pad = np.zeros((y.shape[0] + 2 * k, y.shape[1] + 2 * k), dtype=y.dtype)
        pad[k:-k, k:-k] = y
        n0, n1 = y.shape
        out = w[k] * y
        for s in range(1, k + 1):
            out = out + w[k + s] * (pad[k + s:k + s + n0, k:k + n1] + pad[k - s:k - s + n0, k:k + n1]
                                    + pad[k:k + n0, k + s:k + s + n1] + pad[k:k + n0, k - s:k - s + n1])
        return out

    def rhs(y, P):
        usq = y.real ** 2 + y.imag ** 2
        res = c[11] * P / (c[12] + c[13] * usq)
        a = c[2] * res - c[3]
        b = c[4] * usq + c[5] * res
        return (a - 1j * b) * y + 1j * lap(y)

    def step(psi_in, psi_out, pumping, row0, row1):
        u = psi_in.numpy()
        P = pumping.numpy()
        grow = np.arange(u.shape[0]) + plan.global_row0
        mask = ((grow >= 0) & (grow < plan.n)).astype(float)[:, None]   # rows outside the square stay zero
        k1 = rhs(u, P)
        k2 = rhs(mask * (u + (dt / 2) * k1), P)
        k3 = rhs(mask * (u + (dt / 2) * k2), P)
        k4 = rhs(mask * (u + dt * k3), P)
        new = mask * (u + (dt / 6) * (k1 + 2 * k2 + 2 * k3 + k4))
        psi_out.numpy()[row0:row1] = new[row0:row1]

    return step
