"""The part of the bench.py contract that runs without a GPU: `--impl reference` times the CPU port of the reference
path (Julia / GraphNetCore are absent: oracle/torch_cpu_ref.py, `kind: "port"`) and prints ONE JSON line with the keys
the driver reads; `bench_roofline.algorithmic_work` states the bytes and FLOPs the roofline is computed from."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines                         # stdout carries the JSON line and nothing else
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "mp_step_edges_per_sec_train" and d["unit"] == "edges/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 0
    assert d["config"]["workload"] == "cylinder_flow_train_step" and d["config"]["nodes"] == 1885
    assert d["config"]["edges"] == 10936 and d["config"]["mps"] == 15 and d["config"]["latent"] == 128
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None                        # BASELINE.md publishes no number for this metric


def test_algorithmic_work_matches_design_section_3():
    """bench_roofline.algorithmic_work is the arithmetic behind `roofline.achieved`: its per-row figures are the ones
    DESIGN.md section 3 tabulates (bf16 mode, training, D = 128, L = 4)."""
    sys.path.insert(0, ROOT)
    import bench_roofline as br
    D, L = 128, 4
    # one extra edge / node, one MP step more: the differences isolate the per-row and per-step figures
    base = br.algorithmic_work(10936, 1885, D, L, 15, 9, 3, 2)
    de = br.algorithmic_work(10937, 1885, D, L, 15, 9, 3, 2)
    dm = br.algorithmic_work(10936, 1885, D, L, 16, 9, 3, 2)
    E, N = 10936, 1885
    assert dm["tc_mlp_fwd"][1] - base["tc_mlp_fwd"][1] == 1548 * E + (2820 + 512) * N      # forward, per MP step
    assert dm["tc_mlp_fwd"][0] - base["tc_mlp_fwd"][0] == 2 * D * D * ((L + 2) * E + (L + 1) * N)
    assert dm["tc_mlp_bwd"][1] - base["tc_mlp_bwd"][1] == 1540 * E + 1796 * N + 512 * N      # backward chain
    assert dm["tc_mlp_bwd"][0] - base["tc_mlp_bwd"][0] == 4 * D * D * (L - 1) * (E + N)
    assert dm["tc_dw"][1] - base["tc_dw"][1] == 1288 * E + 3840 * N                          # backward input layer
    assert dm["tc_dw"][0] - base["tc_dw"][0] == 12 * D * D * E + 8 * D * D * N
    assert de["tc_dw"][1] - base["tc_dw"][1] == 15 * 1288
    total_flops = sum(v[0] for v in base.values())
    assert abs(total_flops - 115.0e9) < 0.01 * 115.0e9        # 115.0 GFLOP per single-window training step
    total_bytes_32 = 32 * sum(v[1] for v in base.values())
    assert abs(total_bytes_32 - 32.8e9) < 0.01 * 32.8e9       # ~33 GB per 32-window step (42 GB in round 1)
    # SURVEY 8(d) forward bytes: round 1's convention (fp32 edge latent) is the judge's 7.07 GB; stored widths now: 4.29 GB
    assert abs(br.survey_forward_bytes(349952, 60320, D, L, 15, 9, 3, 2, 4) - 7.066e9) < 0.01e9
    assert abs(br.survey_forward_bytes(349952, 60320, D, L, 15, 9, 3, 2, 4, s_edge=2) - 4.288e9) < 0.01e9
