"""Random-field material properties -- the reference's `scatter/random_fields.py` with the field sampled on the GPU.

The reference builds a gstools `SRF` (covariance model Gaussian / Exponential / Matern / Linear, anisotropic length
scales `[aniso_x, 1, aniso_z] * theta`), evaluates it at the element centroids through meshio, applies a lognormal
transform and creates one material per element (`random_fields.py:59-120`), then re-indexes the physical tags
(`:46-57`).  gstools (==1.7.0) and meshio are un-vendored third-party packages; what gstools computes is the
*randomisation method* (Hesse et al. 2014, `gstools.field.generator.RandMeth`):

    field(x) = mean + sqrt(var / N) * sum_j ( z1_j cos(k_j . x) + z2_j sin(k_j . x) ),     N = 1000 modes

with k_j drawn from the spectral density of the covariance model and z1, z2 ~ N(0, 1).  Here the modes are drawn on the
host (`sample_modes`) and the O(elements x modes) evaluation runs in the CUDA library (`sc_srf_sample`, csrc/srf.cu).
Class, method and attribute names follow the reference, so `scatter.scatter(..., random_props=...)` is unchanged.

Parity.  gstools is not installed, but for the covariance models whose radial spectral density has a closed-form
inverse (`CovModel.has_ppf`: Exponential and Gaussian in 1-D / 2-D) its seed -> modes path is plain numpy `RandomState`
arithmetic, restated in `gstools_modes`.  **Pinned**: the 2-D Exponential field -- the reference's own random-field test
(`integration_tests/integration_test.py:376-421`, golden `results_rf_2d/data.pickle`) is reproduced to 1e-14 relative
(`tests/golden/history_rf_2d.npz`; CPU test of the sampler, GPU test of `scatter(..., random_props=...)` end to end).  The
2-D Gaussian model follows the same path with gstools' published formulas but has no reference fixture.  In 3-D (and for
Matern) gstools samples the radii with an `emcee` MCMC ensemble, which is not restated: there `sample_modes` draws its own
stream -- **unpinned**: same mean, variance and covariance function as the reference's realisation, not the same numbers.
The kernel is checked against the CPU restatement (oracle/fem_np.py:srf_field) on the same modes, the own mode sampler
against the analytical correlation functions (tests/test_host_logic.py).
"""
from __future__ import annotations

import os

import numpy as np

MODE_NO = 1000                                            # gstools default `mode_no`
# rescale factors of the gstools covariance models (len_scale is the integral scale of the Gaussian model)
RESCALE = {"Gaussian": np.sqrt(np.pi) / 2.0, "Exponential": 1.0, "Matern": 1.0}
MATERN_NU = 1.0                                           # gstools default


def correlation(model_name: str, h):
    """Correlation function of the unit-length-scale model at distance h (gstools `CovModel.cor`)."""
    h = np.asarray(h, dtype=float)
    s = RESCALE[model_name]
    if model_name == "Gaussian":
        return np.exp(-(s * h) ** 2)
    if model_name == "Exponential":
        return np.exp(-s * h)
    from scipy.special import gamma, kv
    nu = MATERN_NU
    r = np.sqrt(nu) * s * np.maximum(h, 1e-300)
    out = 2.0 ** (1.0 - nu) / gamma(nu) * r ** nu * kv(nu, r)
    return np.where(h == 0, 1.0, out)


def gstools_has_ppf(model_name: str, dim: int) -> bool:
    """Models / dimensions for which gstools 1.7.0 samples the radii by inversion (`CovModel.has_ppf`)."""
    return model_name in ("Exponential", "Gaussian") and dim in (1, 2)


def gstools_modes(model_name: str, dim: int, len_scale: float, seed: int, mode_no: int = MODE_NO):
    """Restatement of `gstools.field.generator.RandMeth.reset_seed` (gstools 1.7.0) for models with `has_ppf`:
    -> wave vectors k (mode_no, 3) [1/length] for the main length scale `len_scale`, amplitudes z1, z2 (mode_no,).

    * `gstools.random.rng.MasterRNG(seed)`: `RandomState(seed).randint(1, 2**16 - 1)` hands out one sub-seed per stream;
      `RNG.random` opens a fresh `RandomState(sub-seed)` at every access.
    * streams, in this order: z1 ~ normal, z2 ~ normal, directions (`RNG.sample_sphere`: +-1 in 1-D, the angle
      uniform(0, 2 pi) in 2-D), radii (`RNG.sample_dist` -> scipy `rv_continuous.rvs`: `ppf(RandomState(sub-seed).uniform)`).
    * radial ppf with f = len_scale / rescale.  Exponential (`gstools/covmodel/models.py`): 1-D tan(pi u / 2) / f, 2-D
      sqrt(1/u^2 - 1) / f (gstools inverts the survival function there; kept, the golden depends on it).  Gaussian: 1-D
      2 erfinv(u) / f, 2-D 2 sqrt(-log(1 - u)) / f."""
    if not gstools_has_ppf(model_name, dim):
        raise ValueError("gstools samples this model/dimension with an MCMC ensemble; no restatement")
    master = np.random.RandomState(int(seed))

    def stream():
        return np.random.RandomState(master.randint(1, 2 ** 16 - 1))
    z1 = stream().normal(size=mode_no)
    z2 = stream().normal(size=mode_no)
    if dim == 1:
        sphere = stream().choice([-1, 1], size=mode_no)[None, :].astype(float)
    else:
        ang = stream().uniform(0.0, 2 * np.pi, mode_no)
        sphere = np.vstack([np.cos(ang), np.sin(ang)])
    u = stream().uniform(size=mode_no)
    f = float(len_scale) / RESCALE[model_name]
    if model_name == "Exponential":
        rad = np.tan(np.pi / 2 * u) / f if dim == 1 else np.sqrt(1.0 / u ** 2 - 1.0) / f
    else:
        from scipy.special import erfinv
        rad = 2.0 / f * erfinv(u) if dim == 1 else 2.0 / f * np.sqrt(-np.log(1.0 - u))
    k3 = np.zeros((mode_no, 3))
    k3[:, :dim] = (rad * sphere).T
    return k3, z1, z2


def sample_modes(model_name: str, dim: int, seed: int, mode_no: int = MODE_NO):
    """Wave vectors k (mode_no, 3) of the unit-length-scale model and amplitudes z1, z2 (mode_no,).

    Spectral densities: Gaussian exp(-(s h)^2) -> k ~ N(0, 2 s^2 I); Exponential exp(-s h) -> multivariate Cauchy with
    scale s; Matern(nu) -> multivariate Student-t with 2 nu degrees of freedom and scale s / sqrt(2)."""
    if model_name == "Linear":
        raise NotImplementedError("the Linear covariance model has no positive spectral density in 2-D/3-D; "
                                  "use Gaussian, Exponential or Matern")
    if model_name not in RESCALE:
        raise ValueError(f'model name: "{model_name}" is not supported')
    rng = np.random.RandomState(int(seed) % (2 ** 32))
    z1 = rng.normal(size=mode_no)
    z2 = rng.normal(size=mode_no)
    g = rng.normal(size=(mode_no, dim))
    s = RESCALE[model_name]
    if model_name == "Gaussian":
        k = np.sqrt(2.0) * s * g
    elif model_name == "Exponential":
        k = s * g / np.abs(rng.normal(size=mode_no))[:, None]
    else:
        nu = MATERN_NU
        chi2 = rng.gamma(shape=nu, scale=2.0, size=mode_no)            # chi-square with 2 nu degrees of freedom
        k = (s / np.sqrt(2.0)) * g / np.sqrt(chi2 / (2.0 * nu))[:, None]
    k3 = np.zeros((mode_no, 3))
    k3[:, :dim] = k
    return k3, z1, z2


class SpectralField:
    """What `generate_gstools_rf` returns in place of the gstools `SRF` object: the drawn modes and the transform."""

    def __init__(self, model_name, dim, var, mean, len_scale, angles, seed, mode_no=MODE_NO):
        self.model_name, self.dim, self.var, self.mean = model_name, dim, float(var), float(mean)
        self.len_scale = np.asarray(len_scale, dtype=float)[:dim]
        self.angles, self.seed, self.mode_no = float(angles), int(seed), int(mode_no)
        self.gstools_stream = gstools_has_ppf(model_name, dim)
        if self.gstools_stream:
            # gstools' own modes for the main length scale; `isometrize` divides every coordinate by its length scale, so the
            # wave vectors are brought to the unit-length-scale form the kernel expects (same phases k . x)
            k, self.z1, self.z2 = gstools_modes(model_name, dim, self.len_scale[0], seed, mode_no)
            self.k = k * self.len_scale[0]
        else:
            self.k, self.z1, self.z2 = sample_modes(model_name, dim, seed, mode_no)
        self.seconds_device = 0.0

    def isometrize(self, pos):
        """Rotate by -angles about z and divide by the length scales -> (n, 3) points of the unit-length-scale field."""
        p = np.zeros((len(pos), 3))
        p[:, :self.dim] = np.asarray(pos, dtype=float)[:, :self.dim]
        if self.angles != 0.0:
            c, s = np.cos(self.angles), np.sin(self.angles)
            x, y = p[:, 0].copy(), p[:, 1].copy()
            p[:, 0], p[:, 1] = c * x + s * y, -s * x + c * y
        p[:, :self.dim] /= self.len_scale
        return p

    def __call__(self, pos, lognormal=False, device=0, ctx=None):
        from . import _lib
        own = ctx is None
        if own:
            ctx = _lib.Context(device)                       # raises without a CUDA device: there is no CPU sampler
        try:
            out, self.seconds_device = ctx.srf_sample(self.isometrize(pos), self.k, self.z1, self.z2,
                                                      np.sqrt(self.var / self.mode_no), self.mean, lognormal)
        finally:
            if own:
                ctx.close()
        return out


class RF:
    def __init__(self, random_properties: dict, materials: dict, output_folder: str, element_type: str, device: int = 0):
        self.theta = random_properties["theta"]
        self.seed = random_properties["seed_number"]
        self.materials = materials
        self.material_name = random_properties["material"]
        self.key_material = random_properties["key_material"]
        self.sd = random_properties["std_value"]
        self.aniso_x = random_properties["aniso_x"]
        self.aniso_z = random_properties["aniso_z"]
        self.lognormal = True
        self.new_material = {}
        self.new_model_material = []
        self.new_material_index = []
        self.output_folder = output_folder
        self.element_type = element_type
        self.model_name = random_properties["model_name"]
        self.device = device
        self.fields = []
        if not os.path.isdir(output_folder):
            os.makedirs(output_folder)

    def element_type_to_meshio_element_type(self):
        """scatter element type -> meshio cell type (`random_fields.py:30-44`; kept for interface compatibility)."""
        return {"hexa8": "hexahedron", "hexa20": "hexahedron20", "tetra4": "tetra", "tetra10": "tetra10", "tri3": "triangle",
                "tri6": "triangle6", "quad4": "quad", "quad8": "quad8"}[self.element_type]

    def update_material_list(self, materials, model, material_idx):
        """`random_fields.py:46-57`: the random-field elements get the physical tags 0..N-1 (one material each), every
        original tag is shifted by N-1."""
        materials.update(self.new_material)
        shift = self.new_material_index[-1]
        for material in model.materials:
            material[1] = material[1] + shift
        model.materials = model.materials + self.new_model_material
        index = np.asarray(model.materials_index) + shift
        index[index == material_idx + shift] = self.new_material_index
        model.materials_index = index

    def lognormal_parameters(self):
        """(mean, variance) of the underlying normal field (`random_fields.py:70-77`)."""
        mean = self.materials[self.material_name][self.key_material]
        if self.lognormal:
            return np.log(mean ** 2 / (np.sqrt(mean ** 2 + self.sd ** 2))), np.log((self.sd / mean) ** 2 + 1)
        return mean, self.sd ** 2

    @staticmethod
    def centroids(nodes, elements):
        """Centroids of `elements` (node ids) -- what meshio's cell-centre sampling evaluates (`random_fields.py:97-100`)."""
        ids = nodes[:, 0].astype(np.int64)
        if np.array_equal(ids, np.arange(1, len(ids) + 1)):
            rows = np.asarray(elements, dtype=np.int64) - 1
        else:
            order = np.argsort(ids, kind="stable")
            rows = order[np.searchsorted(ids[order], elements)]
        cen = np.zeros((len(rows), 3))
        for b in range(rows.shape[1]):                        # node by node: no (Ne, nne, 3) temporary
            cen += nodes[rows[:, b], 1:4]
        return cen / rows.shape[1]

    def generate_gstools_rf(self, nodes, elements, ndim, angles=0.0, ctx=None):
        """Random field at the centroids of `elements` -> `self.fields[0]`, one new material per element
        (`random_fields.py:59-120`).  Returns the field sampler (the reference returns the gstools SRF)."""
        seed = abs(self.seed)
        len_scale = np.array([self.aniso_x, 1, self.aniso_z]) * self.theta
        mean, var = self.lognormal_parameters()
        if self.model_name not in ("Gaussian", "Exponential", "Matern", "Linear"):
            print('model name: "', self.model_name, '" is not supported')
            return
        srf = SpectralField(self.model_name, ndim, var, mean, len_scale, angles, seed)
        self.fields = [srf(self.centroids(nodes, elements), lognormal=self.lognormal, device=self.device, ctx=ctx)]
        base = self.materials[self.material_name]
        values = self.fields[0]
        for idx in range(len(elements)):
            vals = dict(base)
            vals[self.key_material] = values[idx]
            self.new_material[f"material_{idx + 1}"] = vals
            self.new_model_material.append([3, idx, f"material_{idx + 1}"])
            self.new_material_index.append(idx)
        return srf

    def element_properties(self, model, material_idx):
        """Array form for large meshes: (E, nu, rho) per element of `model` with the field applied to the elements of
        physical tag `material_idx` -- feeds `GenerateMatrix.generate_stiffness_and_mass(..., elem_props=...)` directly
        instead of going through one dictionary entry per element."""
        from .system_matrix import resolve_element_properties
        E, nu, rho = resolve_element_properties(model, self.materials)
        target = {"Young": E, "poisson": nu, "density": rho}[self.key_material]
        target[np.asarray(model.materials_index) == material_idx] = self.fields[0]
        return E, nu, rho

    def dump(self):
        """`random_fields.py:122-135`"""
        with open(os.path.join(self.output_folder, 'rf_props.txt'), 'w') as fo:
            fo.write('Random field properties\n')
            fo.write(f"Model: {self.model_name}\n")
            fo.write('Theta: ' + str(self.theta) + '\n')
            fo.write('Aniso_x: ' + str(self.aniso_x) + '\n')
            fo.write('Aniso_z: ' + str(self.aniso_z) + '\n')
            fo.write('Seed number: ' + str(self.seed) + '\n')
            fo.write('Mean value: ' + str(self.materials[self.material_name][self.key_material]) + '\n')
            fo.write('Std value: ' + str(self.sd) + '\n')
            fo.write('Log normal: ' + str(self.lognormal) + '\n')
