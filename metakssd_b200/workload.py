"""Synthetic workload construction for bench.py / smoke, using only the product
library: reads and genomes are generated on the device (include/mkssd_synth.h), the MarkerDB is
built from the genomes' FASTA sketches with the semantics of the reference's
`set -g` / `set -q` / `set -i` pipeline (command_set.c:831, 427, 322): with one genome per
species, a species keeps exactly the codes no other species has.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .api import Sketcher, SynthSpec, make_shuf, synth_spec


@dataclass
class MarkerDB:
    names: list                 # "<taxid>_<name>" like grouping_genomes() writes them
    comp: list                  # per component: (codes uint32[], index uint64[S+1])

    @property
    def n_codes(self) -> int:
        return int(sum(c[0].size for c in self.comp))


def markerdb_from_species_sketches(sketches, component_num: int) -> list:
    """sketches[s].codes[c] -> per component (codes, index) keeping codes unique to one species."""
    out = []
    S = len(sketches)
    for c in range(component_num):
        per = [sk.codes[c] for sk in sketches]
        sizes = np.array([p.size for p in per], dtype=np.int64)
        allc = np.concatenate(per) if S else np.empty(0, np.uint32)
        owner = np.repeat(np.arange(S, dtype=np.int64), sizes)
        # a code is a marker iff it occurs in exactly one species' sketch
        uniq, inv, cnt = np.unique(allc, return_inverse=True, return_counts=True)
        keep = cnt[inv] == 1
        codes = allc[keep]
        kept_owner = owner[keep]
        index = np.zeros(S + 1, dtype=np.uint64)
        index[1:] = np.cumsum(np.bincount(kept_owner, minlength=S))
        out.append((codes.astype(np.uint32), index))
    return out


def build_markerdb(sk: Sketcher, spec: SynthSpec, batch_bytes: int = 1 << 30) -> MarkerDB:
    """Generate every species genome on the device, sketch it (FASTA path), keep unique codes."""
    import torch

    S = int(spec.P.n_species)
    per_file = spec.fasta_bytes(0) + 16
    per_batch = max(1, min(S, batch_bytes // per_file))
    buf = torch.empty(per_batch * per_file + 256, dtype=torch.uint8, device="cuda:%d" % sk.info.device)
    sketches = []
    for s0 in range(0, S, per_batch):
        s1 = min(S, s0 + per_batch)
        off = sk.synth_fasta_device(spec.P, s0, s1, buf, buf.numel())
        sketches += sk.fasta_co_device(buf, off)
    del buf
    names = ["%d_sp%d" % (s + 1, s) for s in range(S)]
    return MarkerDB(names, markerdb_from_species_sketches(sketches, sk.info.component_num))


def split_by_code_range(codes: np.ndarray, n_parts: int, code_bits: int):
    """Boundaries of n_parts equal ranges of the code space [0, 2^code_bits)."""
    edges = [(i << code_bits) // n_parts for i in range(n_parts + 1)]
    return np.searchsorted(codes, np.array(edges, dtype=np.uint64), side="left")


__all__ = ["MarkerDB", "build_markerdb", "markerdb_from_species_sketches", "split_by_code_range", "make_shuf",
           "synth_spec"]
