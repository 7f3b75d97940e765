"""Host-side path set-up that feeds the hot path (row N1 of SURVEY §8f): natural cubic splines
(`spline`/`tridag`/`splint`/`splin_grad`/`locate`, instantonmod.f90:397-596), the reaction-coordinate
normalisation of read_path (instantonmod.f90:923-936) and the Gauss-Legendre end points / tangents
of pimd_par.f90:212-221.  O(npath) work done once per run; plain numpy."""
import numpy as np


def spline(x, y, yp1=1.0e31, ypn=1.0e31):
    """second derivatives y2 of the interpolating cubic (natural when yp > 0.99e30)"""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n = x.size
    a = np.zeros(n + 1)
    b = np.zeros(n + 1)
    c = np.zeros(n + 1)
    r = np.zeros(n + 1)
    c[1:n] = x[1:] - x[:-1]
    r[1:n] = 6.0 * ((y[1:] - y[:-1]) / c[1:n])
    r[2:n] = r[2:n] - r[1:n - 1]
    a[2:n] = c[1:n - 1]
    b[2:n] = 2.0 * (c[2:n] + a[2:n])
    b[1] = b[n] = 1.0
    if yp1 > 0.99e30:
        r[1] = c[1] = 0.0
    else:
        r[1] = (3.0 / (x[1] - x[0])) * ((y[1] - y[0]) / (x[1] - x[0]) - yp1)
        c[1] = 0.5
    if ypn > 0.99e30:
        r[n] = a[n] = 0.0
    else:
        r[n] = (-3.0 / (x[-1] - x[-2])) * ((y[-1] - y[-2]) / (x[-1] - x[-2]) - ypn)
        a[n] = 0.5
    u = np.zeros(n + 1)
    gam = np.zeros(n + 2)
    bet = b[1]
    u[1] = r[1] / bet
    for j in range(2, n + 1):
        gam[j] = c[j - 1] / bet
        bet = b[j] - a[j] * gam[j]
        if bet == 0.0:
            raise ZeroDivisionError("tridag_ser: Error at code stage 2")
        u[j] = (r[j] - a[j] * u[j - 1]) / bet
    for j in range(n - 1, 0, -1):
        u[j] = u[j] - gam[j + 1] * u[j + 1]
    return u[1:].copy()


def locate(xx, x):
    n = len(xx)
    ascnd = xx[-1] >= xx[0]
    jl, ju = 0, n + 1
    while ju - jl > 1:
        jm = (ju + jl) // 2
        if ascnd == (x >= xx[jm - 1]):
            jl = jm
        else:
            ju = jm
    if x == xx[0]:
        return 1
    if x == xx[-1]:
        return n - 1
    return jl


def _bracket(xa, x):
    n = len(xa)
    klo = max(min(locate(xa, x), n - 1), 1)
    khi = klo + 1
    h = xa[khi - 1] - xa[klo - 1]
    if h == 0.0:
        raise ZeroDivisionError("bad xa input in splint")
    return klo, khi, h, (xa[khi - 1] - x) / h, (x - xa[klo - 1]) / h


def splint(xa, ya, y2a, x):
    klo, khi, h, a, b = _bracket(xa, x)
    return a * ya[klo - 1] + b * ya[khi - 1] + ((a ** 3 - a) * y2a[klo - 1] + (b ** 3 - b) * y2a[khi - 1]) * (h ** 2) / 6.0


def splin_grad(xa, ya, y2a, x):
    klo, khi, h, a, b = _bracket(xa, x)
    return ((ya[khi - 1] - ya[klo - 1]) / h) + ((1.0 - 3.0 * a ** 2) * y2a[klo - 1] + (3.0 * b ** 2 - 1.0) * y2a[khi - 1]) * h / 6.0


def build_path(points):
    """points(npath, ndim, natom) -> (lampath, path, splinepath) as read_path leaves them:
    lampath = cumulative Euclidean distance normalised to [0,1]; natural splines per coordinate."""
    path = np.asfortranarray(points, dtype=np.float64)
    npath = path.shape[0]
    lam = np.zeros(npath)
    for i in range(1, npath):
        lam[i] = lam[i - 1] + np.sqrt(np.sum((path[i] - path[i - 1]) ** 2))
    lam = lam / lam[-1]
    spl = np.zeros_like(path, order="F")
    for i in range(path.shape[1]):
        for j in range(path.shape[2]):
            spl[:, i, j] = spline(lam, path[:, i, j])
    return lam, path, spl


def endpoints(lam, path, spl, xi):
    """xint(k,:,:) and dbdxi(k,:,:) of pimd_par.f90:214-221 for the quadrature nodes xi"""
    nd, na = path.shape[1], path.shape[2]
    xint = np.empty((nd, na, len(xi)), order="F")
    dbd = np.empty_like(xint)
    for k, x in enumerate(xi):
        for i in range(nd):
            for j in range(na):
                xint[i, j, k] = splint(lam, path[:, i, j], spl[:, i, j], x)
                dbd[i, j, k] = splin_grad(lam, path[:, i, j], spl[:, i, j], x)
    return xint, dbd


def acceptor_switch_path(well1, well2, npath=9):
    """Synthetic acceptor-switch path for the water dimer (atoms O,H,H | O,H,H as columns of (3,6)):
    the acceptor's hydrogens (atoms 5,6) rotate by pi about the acceptor's bisector through O_b, with a
    linear blend so that the last point is exactly well2.  (A straight line well1->well2 would drive the
    two hydrogens through each other at lambda = 1/2.)"""
    w1 = np.asarray(well1, dtype=np.float64)
    w2 = np.asarray(well2, dtype=np.float64)
    O, H1, H2 = w1[:, 3], w1[:, 4], w1[:, 5]
    u = 0.5 * (H1 + H2) - O
    u = u / np.linalg.norm(u)

    def rot(v, th):
        return v * np.cos(th) + np.cross(u, v) * np.sin(th) + u * np.dot(u, v) * (1.0 - np.cos(th))

    end = w1.copy()
    end[:, 4] = O + rot(H1 - O, np.pi)
    end[:, 5] = O + rot(H2 - O, np.pi)
    pts = np.empty((npath, 3, 6), order="F")
    for k in range(npath):
        s = k / (npath - 1)
        g = w1.copy()
        g[:, 4] = O + rot(H1 - O, s * np.pi)
        g[:, 5] = O + rot(H2 - O, s * np.pi)
        pts[k] = g + s * (w2 - end)
    return pts


# ---- alignment of the wells (instantonmod.f90:222-376), host side, O(natom) ---------------------------------------
def rotate_atoms(atoms, axis, theta):
    """instantonmod.f90:346-376: (3, natom) coordinates rotated about `axis` (1, 2, 3) by theta; identity for
    |theta| <= 1e-10 like the reference"""
    atoms = np.array(atoms, dtype=np.float64)
    if abs(theta) <= 1e-10:
        return atoms
    j, k = {1: (1, 2), 2: (0, 2), 3: (0, 1)}[axis]
    r = np.zeros((3, 3))
    r[axis - 1, axis - 1] = 1.0
    r[j, j] = np.cos(theta)
    r[k, k] = np.cos(theta)
    r[k, j] = -np.sin(theta)
    r[j, k] = np.sin(theta)
    return r @ atoms


def get_align(atomsin, atom1=1, atom2=2, atom3=3):
    """get_align (instantonmod.f90:222-264): angles that put atom1 at the origin, atom1->atom2 on the x axis and atom3
    in the xz plane, and the origin shift.  atoms (3, natom); atom indices 1-based like mcmod_mass's atom1..3."""
    atomsin = np.asarray(atomsin, dtype=np.float64)
    if atomsin.shape[0] != 3:
        raise ValueError("Wrong number of dimensions; change align_atoms subroutine!")
    i1, i2, i3 = atom1 - 1, atom2 - 1, atom3 - 1
    origin = atomsin[:, i1].copy()
    atoms = atomsin - origin[:, None]
    w = atoms[:, i2] - atoms[:, i1]
    theta1 = float(np.arctan2(w[1], w[0]))
    atoms = rotate_atoms(atoms, 3, theta1)
    w = atoms[:, i2] - atoms[:, i1]
    theta2 = float(np.arctan2(w[2], w[0]))
    atoms = rotate_atoms(atoms, 2, theta2)
    w = atoms[:, i3] - atoms[:, i1]
    theta3 = float(-np.arctan2(w[1], w[2]))
    return theta1, theta2, theta3, origin


def align_atoms(atomsin, theta1, theta2, theta3, origin=None, atom1=1):
    """align_atoms (instantonmod.f90:269-312): atom1 to the origin (its own position, as in the reference — `origin`
    is accepted and unused there), then the three rotations; the input unchanged if all angles are below 1e-10"""
    atomsin = np.asarray(atomsin, dtype=np.float64)
    if atomsin.shape[0] != 3:
        raise ValueError("Wrong number of dimensions; change align_atoms subroutine!")
    if not any(abs(t) > 1e-10 for t in (theta1, theta2, theta3)):
        return atomsin.copy()
    out = atomsin - atomsin[:, atom1 - 1][:, None]
    out = rotate_atoms(out, 3, theta1)
    out = rotate_atoms(out, 2, theta2)
    return rotate_atoms(out, 1, theta3)


def align_wells(well1, well2, alignwell=False, atoms=(1, 2, 3)):
    """pimd_par.f90:159-165: well1 aligned by its own angles; well2 by the same angles, or by its own if alignwell"""
    t = get_align(well1, *atoms)
    w1 = align_atoms(well1, t[0], t[1], t[2], t[3], atoms[0])
    if alignwell:
        t = get_align(well2, *atoms)
    w2 = align_atoms(well2, t[0], t[1], t[2], t[3], atoms[0])
    return np.asfortranarray(w1), np.asfortranarray(w2)


# ---- read_path (instantonmod.f90:895-1035) -------------------------------------------------------------------------
def read_xyz_frames(filename, ndim, natom, xunit=1):
    """path.xyz: per frame a count line, a comment line, natom lines `label x [y [z]]` (instantonmod.f90:917-926);
    xunit = 2: Angstrom -> bohr by /0.529177.  Returns points(npath, ndim, natom)."""
    frames = []
    with open(filename) as f:
        lines = [l for l in f.read().splitlines()]
    i = 0
    while i < len(lines):
        if not lines[i].strip():
            i += 1
            continue
        int(lines[i].split()[0])            # `read(15,*) dummy`
        rows = lines[i + 2:i + 2 + natom]
        if len(rows) < natom:
            raise ValueError("truncated frame in %s" % filename)
        pt = np.array([[float(t.lower().replace("d", "e")) for t in r.split()[1:1 + ndim]] for r in rows]).T
        frames.append(pt / 0.529177 if xunit == 2 else pt)
        i += 2 + natom
    return np.array(frames)


def findmiddle(x1, x2, lampath, vpath):
    """findmiddle (instantonmod.f90:832-871): bisection (40 steps, 1e-6) for the zero of dV/dlambda of the splined
    V(lambda) between x1 and x2 — the top of the barrier along the path"""
    lampath, vpath = np.asarray(lampath, dtype=np.float64), np.asarray(vpath, dtype=np.float64)
    spl = spline(lampath, vpath)
    fmid = splin_grad(lampath, vpath, spl, x2)
    f = splin_grad(lampath, vpath, spl, x1)
    if f * fmid >= 0.0:
        raise ValueError("root must be bracketed in findmiddle")
    if f < 0.0:
        mid, dx = x1, x2 - x1
    else:
        mid, dx = x2, x1 - x2
    for _ in range(40):
        dx = dx * 0.5
        xm = mid + dx
        fm = splin_grad(lampath, vpath, spl, xm)
        if fm <= 0.0:
            mid = xm
        if abs(dx) < 1e-6 or fm == 0.0:
            return mid
    raise ValueError("too many bisections in findmiddle")


def centre_lampath(lampath, vpath):
    """the `centre` branch of read_path (:996-1010): reparametrise lambda so that the barrier top sits at 1/2"""
    lampath = np.asarray(lampath, dtype=np.float64)
    xmiddle = findmiddle(0.3, 0.7, lampath, vpath)
    a = 2.0 - 4.0 * xmiddle
    b = 4.0 * xmiddle - 1.0
    if a >= 0:
        out = -0.5 * b / a + np.sqrt((lampath / a) + (0.5 * b / a) ** 2)
    else:
        out = -0.5 * b / a - np.sqrt((lampath / a) + (0.5 * b / a) ** 2)
    return out, a, b, xmiddle


def read_path(points, V_batch, n, align=True, instanton=None, well1=None, well2=None, fixedends=True, centre=False,
              atoms=(1, 2, 3)):
    """read_path (instantonmod.f90:895-1035) on frames already parsed (read_xyz_frames):
      * frames aligned by the FIRST frame's angles (:927-931; for ndim != 3 the reference STOPs in get_align — there the
        frames are used as given),
      * lampath = cumulative Euclidean distance / total (:932, 936), Vpath = V along the path (V_batch: callable on
        (ndim, natom, npath) -> energies; the PES plugin's batched V),
      * xtilde(k) = path(lambda = (k-1)/(n-1)) (:947-949),
      * instanton (callable xtilde -> optimised xtilde; InstantonMod.instanton) refines the path (:953-990): new path =
        well1, xtilde, well2 (fixedends) or xtilde, re-parametrised by arc length,
      * centre: barrier top moved to lambda = 1/2 (:996-1010),
      * natural splines of every coordinate (:1014-1022).
    Returns dict(lampath, path, splinepath, Vpath, xtilde[, centre=(a, b, xmiddle)])."""
    pts = np.asarray(points, dtype=np.float64)
    if align and pts.shape[1] == 3:
        t = get_align(pts[0], *atoms)
        pts = np.array([align_atoms(f, t[0], t[1], t[2], t[3], atoms[0]) for f in pts])
    lam, path, spl = build_path(pts)
    nd, na = path.shape[1], path.shape[2]

    def beads(lam, path, spl):
        xt = np.empty((n, nd, na), order="F")
        for i in range(nd):
            for j in range(na):
                for k in range(n):
                    xt[k, i, j] = splint(lam, path[:, i, j], spl[:, i, j], k / (n - 1.0))
        return xt

    xtilde = beads(lam, path, spl)
    if instanton is not None:
        xtilde = np.asfortranarray(instanton(xtilde))
        if fixedends:
            pts = np.concatenate([np.asarray(well1, dtype=np.float64)[None], xtilde, np.asarray(well2, dtype=np.float64)[None]])
        else:
            pts = np.array(xtilde)
        lam, path, spl = build_path(pts)
    vpath = np.asarray(V_batch(np.asfortranarray(np.moveaxis(path, 0, 2))), dtype=np.float64)
    out = {"Vpath": vpath, "xtilde": xtilde}
    if centre:
        lam, a, b, xm = centre_lampath(lam, vpath)
        out["centre"] = (a, b, xm)
        spl = np.zeros_like(path, order="F")
        for i in range(nd):
            for j in range(na):
                spl[:, i, j] = spline(lam, path[:, i, j])
    out.update({"lampath": lam, "path": path, "splinepath": spl})
    return out
