"""-m gpu: the drop-in command line (gpupsat_b200/gpupsat) on the reference's own tests/cnf fixtures and on generated
instances: verdict line, model line, exit code, autolog.txt — the surface of SATSolver/main.cu + Results.cu."""
import os
import subprocess

import numpy as np
import pytest

from gpupsat_b200.instances import check_model, parse_dimacs_text, pigeonhole, random_ksat, to_dimacs
from tests.helpers import golden

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "gpupsat_b200", "gpupsat")
G = golden()
WORD = {"SAT": "SATISFIABLE", "UNSAT": "UNSATISFIABLE"}


def run_cli(args, cwd):
    out = subprocess.run([CLI] + args, capture_output=True, text=True, cwd=cwd, timeout=120)
    return out.returncode, out.stdout


def model_from_stdout(stdout, n_vars):
    lines = stdout.splitlines()
    i = lines.index("SATISFIABLE")
    vals = [int(t) for t in lines[i + 1].split()]
    assert vals[-1] == 0 and [abs(v) for v in vals[:-1]] == list(range(1, n_vars + 1))
    return np.array([1 if v > 0 else 0 for v in vals[:-1]], dtype=np.uint8)


@pytest.mark.parametrize("name", sorted(G["tests_cnf"]))
@pytest.mark.parametrize("mode", [["-b", "1", "-t", "1"], ["-b", "2", "-t", "2"], []])
def test_reference_fixtures(tmp_path, name, mode):
    """config 1: the reference's 6 golden files, sequential (-b 1 -t 1) and parallel: verdict == .expected"""
    rec = G["tests_cnf"][name]
    path = tmp_path / name
    path.write_text(rec["dimacs"])
    rc, out = run_cli([str(path)] + mode, tmp_path)
    assert rc == 0
    assert WORD[rec["expected"]] in out.splitlines()
    if rec["expected"] == "SAT":
        offs, lits, n_vars = parse_dimacs_text(rec["dimacs"])
        assert check_model(offs, lits, model_from_stdout(out, n_vars))
        assert "Solution was verified" in out


@pytest.mark.parametrize("n,m,seed", [(20, 91, 0), (50, 218, 0), (50, 218, 1), (100, 426, 0), (200, 820, 1)])
def test_generated_instances(tmp_path, n, m, seed):
    offs, lits = random_ksat(n, m, seed)
    path = tmp_path / "f.cnf"
    path.write_text(to_dimacs(offs, lits, n))
    rc_seq, out_seq = run_cli([str(path), "-b", "1", "-t", "1"], tmp_path)
    rc_par, out_par = run_cli(["-i", str(path), "-b", "8", "-t", "32", "-l"], tmp_path)
    assert rc_seq == 0 and rc_par == 0
    v_seq = [w for w in ("SATISFIABLE", "UNSATISFIABLE", "UNDEFINED") if w in out_seq.splitlines()]
    v_par = [w for w in ("SATISFIABLE", "UNSATISFIABLE", "UNDEFINED") if w in out_par.splitlines()]
    assert v_seq == v_par and len(v_seq) == 1 and v_seq[0] != "UNDEFINED"
    assert "About to call sequential kernel!" in out_seq and "About to invoke kernel..." in out_par
    assert "Total time on GPU:" in out_par
    if v_par[0] == "SATISFIABLE":
        assert check_model(offs, lits, model_from_stdout(out_par, n))
    log = (tmp_path / "autolog.txt").read_text().strip().split(",")       # file,threads,blocks,ms (main.cu:316-321)
    assert log[0].endswith("f.cnf") and log[1] == "32" and log[2] == "8" and float(log[3]) >= 0


def test_pigeonhole_and_errors(tmp_path):
    offs, lits = pigeonhole(8, 7)
    path = tmp_path / "php.cnf"
    path.write_text(to_dimacs(offs, lits, 56))
    rc, out = run_cli([str(path)], tmp_path)
    assert rc == 0 and "UNSATISFIABLE" in out.splitlines()
    rc, out = run_cli([str(tmp_path / "missing.cnf")], tmp_path)
    assert rc != 0 and "was not found" in out          # exit(-1), main.cu:115-118
    rc, out = run_cli(["--version"], tmp_path)
    assert rc == 0 and "v0.0.1" in out


def test_statistics_surface_and_simple_strategy(tmp_path):
    """-v 2 prints the reference's statistics sections (Statistics/RuntimeStatistics.cu:299-360) from the kernel's phase
    counters; -s simple = SimpleJobChooser (JobsManager/SimpleJobChooser.cu:22-75)"""
    offs, lits = random_ksat(50, 218, 1)
    p = tmp_path / "u.cnf"
    p.write_text(to_dimacs(offs, lits, 50))
    out = subprocess.run([CLI, str(p), "-b", "2", "-t", "2", "-v", "2"], capture_output=True, text=True, cwd=tmp_path)
    assert out.returncode == 0
    t = out.stdout
    titles = ["*****Statistics******", "Total job's time:", "Pre-processing time:", "Decision time:", "Conflict analyzing time:",
              "Backtracking time:", "Structures reset time:", "Creating structures time:", "Next job time:",
              "Add jobs to assumptions time:", "Processing results time:", "Average backtracked levels:",
              "Pre-processing - handling assumptions time:", "Pre-processing - adding assumptions to graph time:",
              "Pre-processing - adding handling vars time:"]
    pos = [t.index(x) for x in titles]                      # all present ...
    assert pos == sorted(pos)                               # ... in the reference's order
    assert "run " in t.split("Decision time:")[1].split("Conflict analyzing time:")[0]      # decisions were timed
    assert "UNSATISFIABLE" in t
    out = subprocess.run([CLI, str(p), "-s", "simple"], capture_output=True, text=True, cwd=tmp_path)
    assert "Simple jobs generation is ON" in out.stdout and "Number of jobs = 128" in out.stdout and "UNSATISFIABLE" in out.stdout


def test_edge_case_files(tmp_path):
    (tmp_path / "e.cnf").write_text("p cnf 0 0\n")
    out = subprocess.run([CLI, str(tmp_path / "e.cnf")], capture_output=True, text=True, cwd=tmp_path)
    assert out.returncode == 0 and "SATISFIABLE" in out.stdout
    (tmp_path / "f.cnf").write_text("p cnf 3 2\n1 2 3 0\n0\n")
    out = subprocess.run([CLI, str(tmp_path / "f.cnf")], capture_output=True, text=True, cwd=tmp_path)
    assert out.returncode == 0 and "UNSATISFIABLE" in out.stdout
    (tmp_path / "g.cnf").write_text("p cnf 3 3\n1 2 0\n-1 2 0\n3 0\n")          # one live variable pair after the unit
    out = subprocess.run([CLI, str(tmp_path / "g.cnf"), "-b", "4", "-t", "4"], capture_output=True, text=True, cwd=tmp_path)
    assert out.returncode == 0 and "SATISFIABLE" in out.stdout and "was verified" in out.stdout


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_cli_on_all_gpus_gives_the_same_verdicts(tmp_path):
    """gpupsat FILE.cnf -b 64 -t 32 --gpus 0: the C++ multi-GPU host (gpsat_multi_*) behind the drop-in CLI"""
    for n, m, seed, want in ((250, 1065, 0, "UNSATISFIABLE"), (200, 820, 1, "SATISFIABLE")):
        offs, lits = random_ksat(n, m, seed)
        p = tmp_path / f"m{seed}.cnf"
        p.write_text(to_dimacs(offs, lits, n))
        one = subprocess.run([CLI, str(p), "-b", "64", "-t", "32"], capture_output=True, text=True, cwd=tmp_path)
        many = subprocess.run([CLI, str(p), "-b", "64", "-t", "32", "--gpus", "0", "-v", "2"], capture_output=True, text=True, cwd=tmp_path)
        assert one.returncode == 0 and many.returncode == 0, many.stdout[-500:]
        assert want in one.stdout.splitlines() and want in many.stdout.splitlines()
        assert "Number of GPUs:" in many.stdout
        if want == "SATISFIABLE":
            assert "Solution was verified" in many.stdout
