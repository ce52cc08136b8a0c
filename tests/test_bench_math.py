"""bench.py's roofline numerators against SURVEY.md section 8(d), and its hard-coded ncu traffic
figures against the capture committed under profiles/ (CPU only; no kernel runs)."""
import csv
import os

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGE_OF = (("pose_bin_kernel", "pose_scatter"), ("blurz_drc_fwd_kernel", "blurz_drc_fwd"),
            ("drc_blurz_bwd_fast_kernel", "drc_blurz_bwd"), ("gather_pose_bwd_kernel", "gather_pose_bwd"))


def test_contract_bytes_match_the_survey():
    """A 15 321 600 B and B 118 854 656 B per projection; the stage terms add up to
    the formula (14 G + 72 N + 4 I)."""
    wa, wb = bench.WORKLOADS["A"], bench.WORKLOADS["B"]
    assert bench.algorithmic_bytes(wa["N"], wa["V"], wa["V"]) == 15321600
    assert bench.algorithmic_bytes(wb["N"], wb["V"], wb["V"]) == 118854656
    for w in (wa, wb, bench.WORKLOADS["C3"], bench.WORKLOADS["C5"]):
        stages = bench.stage_algorithmic_bytes(w["N"], w["V"], w["V"])
        assert sum(stages.values()) == bench.algorithmic_bytes(w["N"], w["V"], w["V"])
    # the workloads are BASELINE.json's configurations
    assert (wa["P"], wa["N"], wa["V"], wa["K"], wa["sigma"]) == (64, 8000, 64, 21, 3.0)
    assert (wb["P"], wb["N"], wb["V"], wb["K"]) == (128, 16000, 128, 21)
    assert bench.WORKLOADS["C3"]["P"] == 16 * 4 * 4 and bench.WORKLOADS["C5"]["deterministic"]


def test_fma_counts():
    """2 K FMAs per voxel for the plane kernels' two passes, K + a few for the ray kernels."""
    f = bench.stage_fma(64, 64, 21)
    assert f["blur_xy_fwd"] == f["blur_xy_bwd"] == 2 * 21 * 64 ** 3 == 11010048
    assert 21 * 64 ** 3 < f["blurz_drc_fwd"] < f["drc_blurz_bwd"] < 30 * 64 ** 3
    assert f["pose_scatter"] is None and f["gather_pose_bwd"] is None


def _mbyte(cell):
    value, unit = cell.split()
    return float(value) * {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Gbyte": 1e9}[unit]


def test_traffic_figures_are_the_committed_capture():
    """roofline.traffic is quoted from profiles/r02_ncu_full_step.csv: the numbers in bench.py
    must be that file's (first capture of every kernel), to the three digits they are quoted at."""
    rows = {}
    with open(os.path.join(ROOT, bench.NCU_CAPTURE)) as f:
        for r in csv.DictReader(f):
            name = r["kernel"]
            stage = next((s for k, s in STAGE_OF if k in name), None)
            if stage is None and "blur_xy_kernel" in name:
                # template arguments <V, R, FWD_SCATTER, ...>: the forward kernel scatters
                stage = "blur_xy_fwd" if name.split("<")[1].split(",")[2].strip() == "1" else "blur_xy_bwd"
            if stage and stage not in rows:
                rows[stage] = _mbyte(r["dram_read"]) + _mbyte(r["dram_write"])
    quoted = bench.NCU_TRAFFIC_BYTES["A"]
    assert set(quoted) == set(rows)
    for stage, got in rows.items():
        assert abs(got - quoted[stage]) <= 0.006e6 + 1e-3 * got, (stage, got, quoted[stage])
