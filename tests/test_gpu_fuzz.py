"""GPU: randomised parity of the FASTQ and FASTA paths against the oracle — ragged line structure (blank lines,
missing lines, CR, sequence-looking headers and qualities, reads of 0..3000 bases), tiny tiles (many
tickets and groups per CTA) and forced multi-chunk host uploads.  The generator lives in
tools/fuzz_fastq.py (also usable stand-alone with more cases / other seeds)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [11, 12])
def test_random_fastq_matches_oracle(lib_built, oracle, seed):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_fastq.py"), "120", str(seed)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "cases 120 failures 0" in r.stdout, r.stdout[-2000:]


@pytest.mark.parametrize("seed", [21])
def test_random_fasta_batches_match_oracle(lib_built, oracle, seed):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_fasta.py"), "80", str(seed)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "cases 80 failures 0" in r.stdout, r.stdout[-2000:]
