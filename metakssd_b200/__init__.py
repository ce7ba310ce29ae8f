"""metakssd_b200 — B200-native (sm_100a) implementation of MetaKSSD's hot path:
FASTQ -> KSSD sketch with k-mer counts (`dist -L <shuf> -A`) -> `composite` MarkerDB lookup.

The compute lives in libmkssd_b200.so (hand-written CUDA, C ABI in include/mkssd_b200.h);
this package is the Python binding plus the host-side mirror of the reference's file formats.
"""
from .api import (MkError, MkInfo, MksParams, MkProfile, Sketch, Sketcher, SpeciesNames, composite_tsv, coverage_tsv,
                  device_count, load,
                  read_shuf, read_sketch_dir, write_shuf, write_sketch_dir, synth_spec, make_shuf, SynthSpec,
                  EXPORTS, LIB_PATH)

__all__ = ["MkError", "MkInfo", "MksParams", "MkProfile", "Sketch", "Sketcher", "SpeciesNames", "composite_tsv", "coverage_tsv", "device_count",
           "load", "synth_spec", "make_shuf", "SynthSpec", "read_shuf", "read_sketch_dir", "write_shuf", "write_sketch_dir", "EXPORTS", "LIB_PATH"]
