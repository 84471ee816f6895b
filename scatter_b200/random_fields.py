"""Random-field material properties -- input contract of the reference's `scatter/random_fields.py`.

The reference samples a gstools `SRF` at the element centroids, applies a lognormal transform and creates one material
per element (`random_fields.py:59-120, 46-57`).  gstools / meshio are third-party samplers that are not part of the hot
path: the hot path consumes per-element (E, nu, rho) arrays.  This module keeps the `RF` class and
its methods; `generate_gstools_rf` uses gstools when it is importable and otherwise raises a clear error, and
`generate_lognormal` offers a dependency-free, spatially uncorrelated substitute with the same transform.
"""
from __future__ import annotations

import os

import numpy as np


class RF:
    def __init__(self, random_properties: dict, materials: dict, output_folder: str, element_type: str):
        self.theta = random_properties["theta"]
        self.seed_number = random_properties["seed_number"]
        self.material_name = random_properties["material"]
        self.key_material = random_properties["key_material"]
        self.std_value = random_properties["std_value"]
        self.aniso_x = random_properties["aniso_x"]
        self.aniso_z = random_properties["aniso_z"]
        self.model_name = random_properties["model_name"]
        self.materials = materials
        self.mean = materials[self.material_name][self.key_material]
        self.output_folder = output_folder
        self.element_type = element_type
        self.new_material = {}
        self.new_material_index = []
        self.random_field = None

    def _lognormal_parameters(self):
        sig2 = np.log(1.0 + (self.std_value / self.mean) ** 2)       # random_fields.py:73-75
        return np.log(self.mean) - 0.5 * sig2, np.sqrt(sig2)

    def _centroids(self, nodes, elements):
        ids = nodes[:, 0].astype(np.int64)
        order = np.argsort(ids, kind="stable")
        rows = order[np.searchsorted(ids[order], elements)]
        return nodes[rows, 1:].mean(axis=1)

    def generate_gstools_rf(self, nodes, elements, ndim, angles=0.0):
        try:
            import gstools as gs
        except ImportError as exc:
            raise ImportError("random fields with spatial correlation need gstools==1.7.0 (not installed); use "
                              "RF.generate_lognormal or pass per-element arrays") from exc
        cen = self._centroids(nodes, elements)
        model = getattr(gs, self.model_name)
        mu, sigma = self._lognormal_parameters()
        if ndim == 3:
            cov = model(dim=3, var=sigma ** 2, len_scale=self.theta, anis=[self.aniso_x, self.aniso_z], angles=angles)
            pos = [cen[:, 0], cen[:, 1], cen[:, 2]]
        else:
            cov = model(dim=2, var=sigma ** 2, len_scale=self.theta, anis=[self.aniso_x], angles=angles)
            pos = [cen[:, 0], cen[:, 1]]
        srf = gs.SRF(cov, mean=mu, seed=abs(int(self.seed_number)))
        self.random_field = np.exp(srf(pos))

    def generate_lognormal(self, n_elements: int):
        mu, sigma = self._lognormal_parameters()
        rng = np.random.default_rng(abs(int(self.seed_number)))
        self.random_field = np.exp(mu + sigma * rng.standard_normal(n_elements))

    def update_material_list(self, materials, model, material_idx):
        """One new material per random-field element, physical tags re-indexed (random_fields.py:46-57)."""
        existing = [int(m[1]) for m in model.materials]
        next_tag = max(existing) + 1
        sel = np.where(np.asarray(model.materials_index) == material_idx)[0]
        tags = np.asarray(model.materials_index).copy()
        for k, e in enumerate(sel):
            name = f"{self.material_name}_rf_{k}"
            props = dict(materials[self.material_name])
            props[self.key_material] = float(self.random_field[k])
            self.new_material[name] = props
            model.materials.append([float(model.dimension), next_tag + k, name])
            tags[e] = next_tag + k
        model.materials_index = tags

    def dump(self):
        os.makedirs(self.output_folder, exist_ok=True)
        with open(os.path.join(self.output_folder, "rf_props.txt"), "w") as f:
            f.write(f"theta {self.theta}\nseed {self.seed_number}\nmodel {self.model_name}\nstd {self.std_value}\n")
