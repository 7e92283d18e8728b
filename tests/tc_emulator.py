"""CPU emulation of the tensor-core convolution engine's DATA FLOW (csrc/conv_tc.cu), used by the CPU tests to
check the host-side weight packer and the kernel's index arithmetic without a GPU.

The emulator restates, byte for byte, what the CUDA kernel does: the TMA box loads (zero fill outside the
tensor), the shared-memory byte offsets of every MMA's A and B descriptors (``step_desc`` / ``plane_window``
copied from the .cu file), the TMEM column windows, the first-touch / accumulate flags and the N rounded up to
16 ("spill").  Uninitialised shared memory and TMEM start as NaN so that any read of bytes the kernel never
wrote, or any accumulate-before-first-touch, poisons the result.  It deliberately does NOT reuse the Python
packer's own helper tables (step_units / tap_of) -- the point is to cross-check the two descriptions.
"""
from __future__ import annotations

import numpy as np

K3, DOWN, UP, K3T = 0, 1, 2, 3
HX, HY = 10, 18
CHUNK_BYTES = HX * HY * 16
STAGE_BYTES = 2 * CHUNK_BYTES


def pad16(n):
    return (n + 15) & ~15


def step_desc(mode, lone, pp, st):
    """(byte offset, LBO bytes) -- transcription of Steps<KIND>::delta() in conv_tc.cu."""
    if mode == K3:
        if not lone:
            return ((st // 3) * HX + (st % 3)) * 16, CHUNK_BYTES
        if st < 3:
            return (st * HX) * 16, 16
        if st == 3:
            return 2 * 16, HX * 16
        return (2 * HX + 1) * 16, 16
    py, px = pp >> 1, pp & 1
    sy0 = (1 - py) if mode == DOWN else py
    sx0 = (1 - px) if mode == DOWN else px
    if not lone:
        return ((sy0 + st // 2) * HX + (sx0 + st % 2)) * 16, CHUNK_BYTES
    return ((sy0 + st) * HX + sx0) * 16, 16


def plane_window(mode, TZ, zi):
    """(lo, hi, jlo, ft) -- transcription of plane_window() in conv_tc.cu."""
    if mode in (K3, K3T):
        lo, hi = max(zi - 2, 0), min(zi, TZ - 1)
        return lo, hi, 2 - zi + lo, (zi if zi < TZ else TZ)
    if mode == DOWN:
        q = zi >> 1
        lo, hi = max(q - 1, 0), min(q, TZ - 1)
        return lo, hi, lo - (q - 1), (q if (zi & 1) == 0 and q < TZ else TZ)
    lo, hi = max(2 * zi - 3, 0), min(2 * zi, TZ - 1)
    return lo, hi, lo - (2 * zi - 3), max(2 * zi - 1, 0)


def _operand(mem, start_byte, rows, lbo, sbo):
    """Gather a (rows x 16) K-major no-swizzle operand from a flat element array (2 bytes per element)."""
    r = np.arange(rows)[:, None]
    k = np.arange(16)[None, :]
    addr = start_byte + (r // 8) * sbo + (r % 8) * 16 + (k // 8) * lbo + (k % 8) * 2
    return mem[addr // 2]


def emulate_conv_tc(mode, x_blocked, wpacked, cout, TZ=None):
    """x_blocked: float array [N][C8][Z][Y][X][8] (values already bf16-representable);
    wpacked: float array [pass][image][step][2][NB][8].  Returns raw accumulators [N][Cpad][oz][oy][ox]
    (K3T: the tap-summed outputs [N][cout][oz][oy][ox])."""
    N, C8, Z, Y, X, _ = x_blocked.shape
    cpad = ((9 * cout if mode == K3T else cout) + 7) // 8 * 8
    blocks = {K3: 3, DOWN: 2, UP: 4, K3T: 3}[mode]
    NB = blocks * cpad + 16
    G = (C8 + 1) // 2
    steps_full, steps_lone = {K3: (9, 5), K3T: (G, G)}.get(mode, (4, 2))
    lone_last = C8 % 2
    n_pass = 4 if mode == UP else 1
    n_bimg = {DOWN: 8 * G, K3T: 1}.get(mode, G)
    assert wpacked.shape == (n_pass, n_bimg, steps_full, 2, NB, 8), wpacked.shape
    if mode in (K3, K3T):
        oz, oy, ox = Z, Y, X
    elif mode == DOWN:
        oz, oy, ox = Z // 2, Y // 2, X // 2
    else:
        oz, oy, ox = 2 * Z, 2 * Y, 2 * X
    tzmax = min(12, (256 - 16) // cpad)
    if mode == UP:
        tzmax &= ~1
    if TZ is None:
        ntz = -(-oz // tzmax)
        TZ = -(-oz // ntz)
        if mode == UP and TZ & 1:
            TZ += 1
    maxp = min(4, 256 // cpad)
    if cpad % 16 and maxp >= 2:
        maxp &= ~1
    maxp = max(maxp, 1)
    tiles_z = -(-oz // TZ)
    if mode == UP:
        tiles_x, tiles_y, lim_y, lim_x, zin = -(-X // 8), -(-Y // 16), Y, X, TZ // 2 + 2
    else:
        tiles_x, tiles_y, lim_y, lim_x = -(-ox // 8), -(-oy // 16), oy, ox
        zin = TZ + 2 if mode in (K3, K3T) else 2 * TZ + 2
    sx, sy = 8, 16
    if mode == K3T:      # tiles own the 14 x 6 interior of the 16 x 8 voxels they compute tap products for
        sx, sy = 6, 14
        tiles_x, tiles_y = -(-ox // 6), -(-oy // 14)
    out = np.full((N, cout if mode == K3T else cpad, oz, oy, ox), np.nan, dtype=np.float64)

    def load_box(n, chunk, nch, cx, cy, cz, pp):
        """TMA box load -> flat stage array (elements); bytes not written stay NaN."""
        stage = np.full(max(STAGE_BYTES, nch * CHUNK_BYTES) // 2, np.nan)
        box = np.zeros((nch, HY, HX, 8))
        for c in range(nch):
            for yy in range(HY):
                for xx in range(HX):
                    if mode == DOWN:
                        py, px = pp >> 1, pp & 1
                        gx, gy, gz = 2 * (cx + xx) + px, 2 * (cy + yy) + py, cz
                        inside = 0 <= cx + xx < X // 2 and 0 <= cy + yy < Y // 2 and 0 <= gz < Z
                    else:
                        gx, gy, gz = cx + xx, cy + yy, cz
                        inside = 0 <= gx < X and 0 <= gy < Y and 0 <= gz < Z
                    if inside and chunk + c < C8:
                        box[c, yy, xx] = x_blocked[n, chunk + c, gz, gy, gx]
        stage[: nch * CHUNK_BYTES // 2] = box.reshape(-1)
        return stage

    for n in range(N):
        for tz in range(tiles_z):
            for ty in range(tiles_y):
                for tx in range(tiles_x):
                    x0, y0, z0 = tx * sx, ty * sy, tz * TZ
                    tmem = np.full((128, TZ * cpad + 16), np.nan)
                    for ps in range(n_pass):
                        for bi in range(n_bimg):
                            g = bi % G
                            lone = bool(lone_last) and g == G - 1 and mode != K3T
                            nsteps = steps_lone if lone else steps_full
                            bimg = np.full(steps_full * 2 * NB * 8, np.nan)
                            bimg[: nsteps * 2 * NB * 8] = wpacked[ps, bi, :nsteps].reshape(-1)
                            zi_start, zi_step, pp = 0, 1, 0
                            if mode == DOWN:
                                zi_start, zi_step, pp = bi // (4 * G), 2, (bi // G) % 4
                            elif mode == UP:
                                pp = ps
                            for zi in range(zi_start, zin, zi_step):
                                nch = 1 if lone else 2
                                if mode == K3T:      # one box with every chunk of the plane
                                    stage = load_box(n, 0, C8, x0 - 1, y0 - 1, z0 - 1 + zi, 0)
                                elif mode == K3:
                                    stage = load_box(n, 2 * g, nch, x0 - 1, y0 - 1, z0 - 1 + zi, 0)
                                elif mode == UP:
                                    stage = load_box(n, 2 * g, nch, x0 - 1, y0 - 1, z0 // 2 - 1 + zi, 0)
                                else:
                                    stage = load_box(n, 2 * g, nch, x0 - 1, y0 - 1, 2 * z0 - 1 + zi, pp)
                                lo, hi, jlo, ft = plane_window(mode, TZ, zi)
                                for st in range(nsteps):
                                    if mode == K3T:   # step = chunk group st; the last one may be a lone chunk
                                        a_off = st * STAGE_BYTES
                                        a_lbo = 16 if (lone_last and st == G - 1) else CHUNK_BYTES
                                    else:
                                        a_off, a_lbo = step_desc(mode, lone, pp, st)
                                    A = _operand(stage, a_off, 128, a_lbo, HX * 16)
                                    b_step = st * (2 * NB * 16)
                                    split = min(max(ft, lo), hi + 1) if (bi == 0 and st == 0) else hi + 1
                                    q = lo
                                    while q <= hi:
                                        overwrite = q >= split
                                        lim = hi + 1 if overwrite else split
                                        npl = min(maxp, lim - q)
                                        Nn = pad16(npl * cpad)
                                        B = _operand(bimg, b_step + (jlo + (q - lo)) * cpad * 16, Nn, NB * 16, 128)
                                        prod = A @ B.T
                                        cols = slice(q * cpad, q * cpad + Nn)
                                        tmem[:, cols] = prod if overwrite else tmem[:, cols] + prod
                                        q += npl
                        # epilogue of this pass
                        if mode == K3T:
                            # row m = halo voxel (y0-1+my, x0-1+mx); interior voxels sum the nine tap columns of
                            # their nine neighbours (fixed dy, dx order, like the kernel)
                            for m in range(128):
                                my, mx = m >> 3, m & 7
                                yy, xx = y0 - 1 + my, x0 - 1 + mx
                                if not (1 <= my <= 14 and 1 <= mx <= 6 and yy < lim_y and xx < lim_x):
                                    continue
                                for q in range(TZ):
                                    if z0 + q >= oz:
                                        break
                                    for co in range(cout):
                                        acc = 0.0
                                        for dy in range(3):
                                            for dx in range(3):
                                                acc += tmem[(my + dy - 1) * 8 + (mx + dx - 1),
                                                            q * cpad + (dy * 3 + dx) * cout + co]
                                        out[n, co, z0 + q, yy, xx] = acc
                            continue
                        for m in range(128):
                            my, mx = m >> 3, m & 7
                            if mode == UP:
                                yy, xx = 2 * (y0 + my) + (ps >> 1), 2 * (x0 + mx) + (ps & 1)
                                valid = (y0 + my) < lim_y and (x0 + mx) < lim_x
                            else:
                                yy, xx = y0 + my, x0 + mx
                                valid = yy < lim_y and xx < lim_x
                            if not valid:
                                continue
                            for q in range(TZ):
                                if z0 + q >= oz:
                                    break
                                out[n, :, z0 + q, yy, xx] = tmem[m, q * cpad:(q + 1) * cpad]
    return out
