"""Helpers shared by the golden-vector tests: load a case and rebuild its inputs."""
import hashlib
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def make_inputs(seed, world_size, batch, num_classes, emb, label_pool=None):
    """Same generator as tests/golden/make_golden.py::make_inputs (kept in sync by the sha256 pin)."""
    g = torch.Generator().manual_seed(seed)
    feats, labels, weights = [], [], []
    for r in range(world_size):
        f = torch.nn.functional.normalize(torch.randn(batch, emb, generator=g))
        if label_pool is None:
            l = torch.randint(0, num_classes, (batch,), generator=g)
        else:
            l = label_pool[torch.randint(0, len(label_pool), (batch,), generator=g)]
        nl = num_classes // world_size + int(r < num_classes % world_size)
        w = torch.randn(nl, emb, generator=g) * 0.01
        feats.append(f), labels.append(l.long()), weights.append(w)
    return feats, labels, weights


def sha(arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


class Case:
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
        self.cfg = {k[4:]: self.z[k].item() for k in self.z.files if k.startswith("cfg_")}
        W = self.cfg["world_size"]
        if self.cfg["store_inputs"]:
            self.features = [torch.from_numpy(self.z[f"r{r}/features"]) for r in range(W)]
            self.weights = [torch.from_numpy(self.z[f"r{r}/weight"]) for r in range(W)]
            self.labels = [torch.from_numpy(self.z[f"r{r}/labels"]) for r in range(W)]
        else:
            f, l, w = make_inputs(self.cfg["seed"], W, self.cfg["batch"], self.cfg["num_classes"], self.cfg["emb"])
            if sha([t.numpy() for t in f + l + w]) != str(self.z["input_sha256"]):
                raise RuntimeError("seeded inputs no longer reproduce the golden inputs (torch RNG changed?)")
            self.features, self.labels, self.weights = f, l, w

    def has(self, r, step, key):
        return f"r{r}/s{step}/{key}" in self.z.files

    def get(self, r, step, key):
        return self.z[f"r{r}/s{step}/{key}"]


SMALL_CASES = ["w1_sr1_small", "w1_sr1_s30", "w1_sr01", "w1_sr_pos_overflow", "w2_sr1_ragged", "w2_sr03", "w1_arc_small", "w2_arc_sr03"]


def margin_of(pkg, cfg):
    """The margin object of a golden case: pkg.ArcFace for cfg['loss'] == 'arcface', else pkg.CosFace."""
    cls = pkg.ArcFace if cfg.get("loss", "cosface") == "arcface" else pkg.CosFace
    return cls(s=cfg["s"], m=cfg["m"])
