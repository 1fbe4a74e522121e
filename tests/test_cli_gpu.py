"""The phylonium-b200 command line against the unmodified reference binary
(oracle/_ref/phylonium) on the same FASTA files: stdout must be byte-identical."""
import os
import subprocess

import numpy as np
import pytest

import datasets

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "phylonium_b200", "bin", "phylonium-b200")
THEIRS = os.path.join(ROOT, "oracle", "_ref", "phylonium")


def write_fasta(path, contigs, width=70, lower=False, junk=False):
    with open(path, "w") as f:
        for k, c in enumerate(contigs):
            s = c.decode()
            if lower:
                s = s.lower()
            if junk:
                s = s[:50] + "NNNNnn-" + s[50:]
            f.write(f">contig{k} some description\n")
            for i in range(0, len(s), width):
                f.write(s[i : i + width] + "\n")


@pytest.fixture(scope="module")
def fasta_dir(tmp_path_factory):
    if not (os.path.exists(OURS) and os.path.exists(THEIRS)):
        pytest.skip("needs the host binary and the compiled reference")
    d = tmp_path_factory.mktemp("fasta")
    rng = np.random.default_rng(31)
    base = [datasets.random_dna(rng, 20000), datasets.random_dna(rng, 15000), datasets.random_dna(rng, 9000)]
    write_fasta(d / "alpha.fasta", base)
    write_fasta(d / "beta.fa", [datasets.mutate(rng, c, 0.01) for c in base], width=60, lower=True)
    write_fasta(d / "gamma.fas", [datasets.revcomp(datasets.mutate(rng, base[1], 0.03)), datasets.mutate(rng, base[0], 0.03)], junk=True)
    write_fasta(d / "delta.fasta", [datasets.mutate(rng, b"".join(base), 0.05)])
    write_fasta(d / "epsilon.fasta", [datasets.mutate(rng, base[2], 0.02), datasets.mutate(rng, base[0][:12000], 0.02)])
    return d


CASES = [
    ["-r", "alpha.fasta"],
    ["-r", "delta.fasta", "--distance=raw"],
    ["-r", "beta.fa", "--distance=ani"],
    ["-r", "alpha.fasta", "--complete-deletion"],
    ["-r", "gamma.fas", "--2pass"],
    [],
    ["-2"],
]


@pytest.mark.parametrize("opts", CASES, ids=[" ".join(c) or "defaults" for c in CASES])
def test_same_stdout(fasta_dir, opts):
    files = ["alpha.fasta", "beta.fa", "gamma.fas", "delta.fasta", "epsilon.fasta"]
    a = subprocess.run([OURS] + opts + files, cwd=fasta_dir, capture_output=True)
    b = subprocess.run([THEIRS, "-t", "2"] + opts + files, cwd=fasta_dir, capture_output=True)
    assert b.stdout, b.stderr
    assert a.stdout == b.stdout, (a.stderr, b.stderr)
    assert a.returncode == b.returncode


@pytest.mark.parametrize("ref", ["alpha.fasta", "gamma.fas", "delta.fasta"])
def test_reference_positions_file(fasta_dir, ref, tmp_path):
    """-p FILE: the parts of the core genome, their segregating sites and the reference bases
    (src/process.cxx:471-513) — the file and the matrix must be byte-identical"""
    files = ["alpha.fasta", "beta.fa", "gamma.fas", "delta.fasta", "epsilon.fasta"]
    pa, pb_ = tmp_path / "ours.pos", tmp_path / "theirs.pos"
    a = subprocess.run([OURS, "-r", ref, "-p", str(pa)] + files, cwd=fasta_dir, capture_output=True)
    b = subprocess.run([THEIRS, "-t", "2", "-r", ref, "-p", str(pb_)] + files, cwd=fasta_dir, capture_output=True)
    assert b.stdout, b.stderr
    assert a.stdout == b.stdout, (a.stderr, b.stderr)
    assert pb_.read_bytes(), "the reference wrote nothing"
    assert pa.read_bytes() == pb_.read_bytes()
    # an existing file is not overwritten (src/phylonium.cxx:233-240)
    again = subprocess.run([OURS, "-r", ref, "-p", str(pa)] + files, cwd=fasta_dir, capture_output=True)
    assert again.returncode == 1 and b"already exists" in again.stderr


def test_bootstrap_matrices(fasta_dir):
    """-b 5: the matrix itself (byte-identical to the reference's first one) and four
    replicates with resampled substitution counts; the replicates are random in both programs
    (seeded from std::random_device), so only their shape and plausibility are checked"""
    files = ["alpha.fasta", "beta.fa", "gamma.fas", "delta.fasta", "epsilon.fasta"]
    a = subprocess.run([OURS, "-r", "alpha.fasta", "-b", "5"] + files, cwd=fasta_dir, capture_output=True)
    b = subprocess.run([THEIRS, "-t", "2", "-r", "alpha.fasta", "-b", "5"] + files, cwd=fasta_dir, capture_output=True)
    la, lb = a.stdout.decode().splitlines(), b.stdout.decode().splitlines()
    assert len(la) == len(lb) == 5 * 6
    assert la[:6] == lb[:6]

    def matrix(lines):
        return np.array([[float(x) for x in l.split()[1:]] for l in lines[1:6]])

    first = matrix(la[:6])
    for rep in range(1, 5):
        block = la[6 * rep : 6 * rep + 6]
        assert block[0] == "5" and [l.split()[0] for l in block[1:]] == [l.split()[0] for l in la[1:6]]
        m = matrix(block)
        assert (np.diag(m) == 0).all()
        off = ~np.eye(5, dtype=bool)
        assert np.allclose(m[off], first[off], rtol=0.25), rep  # thousands of substitutions per cell
        assert not np.array_equal(m, first)
