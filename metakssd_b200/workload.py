"""Synthetic workload construction for bench.py / smoke, using only the product
library: reads and genomes are generated on the device (include/mkssd_synth.h), the MarkerDB is
built from the genomes' FASTA sketches by the reference's own pipeline run on the device
(`set -g` / `set -q` / `set -i`: mk_set_group, mk_set_uniq_union, mk_set_operate; command_set.c:831, 427, 322).
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


def _next_prime(n: int) -> int:
    while True:
        if all(n % j for j in range(2, int(n ** 0.5) + 1)):
            return n
        n += 1


def organize_taxa(taxids):
    """Order in which organize_taxf() (command_set.c:635-704) lists the taxa: ascending slot of an open-addressing
    table of nextPrime(lines / 0.6) slots.  Returns (taxon position of every genome, taxids in output order)."""
    n = len(taxids)
    hs = _next_prime(int(n / 0.6))
    tab = {}
    slot_of = []
    for t in taxids:
        for i in range(hs):
            hv = (t % hs + i * (1 + t % (hs - 1))) % hs
            if hv not in tab or tab[hv] == t:
                tab[hv] = t
                slot_of.append(hv)
                break
    order = sorted(tab)
    pos = {slot: i for i, slot in enumerate(order)}
    return np.array([pos[s] for s in slot_of], dtype=np.int32), [tab[s] for s in order]


def markerdb_pipeline(sk: Sketcher, sketches, taxids, taxnames) -> MarkerDB:
    """The reference's MarkerDB pipeline on the device: `set -g` (mk_set_group), `set -q` (mk_set_uniq_union),
    `set -i` (mk_set_operate) — species in organize_taxf() order, codes in the order of the per-taxon tables."""
    taxon_of, ids = organize_taxa(taxids)
    name_of = dict(zip(taxids, taxnames))
    comp = []
    for c in range(sk.info.component_num):
        per = [s.codes[c] for s in sketches]
        codes = np.concatenate(per) if per else np.empty(0, np.uint32)
        index = np.zeros(len(per) + 1, dtype=np.uint64)
        index[1:] = np.cumsum([p.size for p in per])
        pan, pan_index = sk.set_group(codes, index, taxon_of, len(ids))
        uniq = sk.set_uniq_union(pan)
        mc, mi = sk.set_operate(uniq, pan, pan_index, True)
        comp.append((mc, mi))
    return MarkerDB(["%d_%s" % (t, name_of[t]) for t in ids], comp)


def build_markerdb(sk: Sketcher, spec: SynthSpec, batch_bytes: int = 1 << 30) -> MarkerDB:
    """Generate every species genome on the device, sketch it (FASTA path), then the set -g / -q / -i pipeline."""
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
    return markerdb_pipeline(sk, sketches, [s + 1 for s in range(S)], ["sp%d" % s for s in range(S)])


__all__ = ["MarkerDB", "build_markerdb", "markerdb_pipeline", "organize_taxa", "make_shuf", "synth_spec"]
