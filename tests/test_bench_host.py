"""Host-side pieces of the measurement chain (no GPU): the roofline bookkeeping of bench.py and the ncu summary tool that
produces the committed profiles it reads."""
import argparse
import importlib.util
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_algorithmic_flops_match_the_survey():
    bench = _load("bench_mod", os.path.join(ROOT, "bench.py"))
    a = argparse.Namespace(workload="tav_roberta_u8", text_len=128)
    assert abs(bench.flop_per_utt(a) - 1555.4e9) / 1555.4e9 < 1e-3            # 1442.9 + 79.1 + 33.35 GFLOP
    a = argparse.Namespace(workload="swin160", text_len=128)
    assert abs(bench.flop_per_utt(a) - 160 * 9.018e9) < 1e6


def test_roofline_traffic_comes_from_the_committed_ncu_pass():
    bench = _load("bench_mod2", os.path.join(ROOT, "bench.py"))
    b, src = bench.ncu_traffic()
    assert b is not None and 1e7 < b < 2e9 and "r02_launches_final.json" in src
    k = bench.ncu_kernel_traffic("mlp_fused C=384 M=125440")
    assert k is not None and 1e8 < k < 2e9
    assert bench.ncu_kernel_traffic("layernorm C=768") is None                # not a GEMM-class kernel of the table
    names = [x["kernel"] for x in json.load(open(os.path.join(ROOT, "profiles", "r02_launches_final.json")))["kernels"]]
    for want in bench.LABEL_TO_NCU.values():
        assert want in names, want


def test_ncu_summary_launch_list(tmp_path):
    tool = _load("ncu_summary", os.path.join(ROOT, "tools", "ncu_summary.py"))
    csv_path = tmp_path / "launches.csv"
    rows = ['==PROF== Connected', '"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size",'
            '"Grid Size","Device","CC","Section Name","Metric Name","Metric Unit","Metric Value"']

    def add(i, name, us, rd_mb, wr_mb):
        base = f'"{i}","1","python","h","{name}","1","7","(128, 1, 1)","(148, 1, 1)","0","10.0","Command line profiler metrics",'
        rows.append(base + f'"gpu__time_duration.sum","us","{us}"')
        rows.append(base + f'"dram__bytes_read.sum","Mbyte","{rd_mb}"')
        rows.append(base + f'"dram__bytes_write.sum","Mbyte","{wr_mb}"')

    add(0, "void fmmt::<unnamed>::warmup_kernel(int)", 5.0, 1, 1)              # skipped (warm-up step)
    add(1, "void fmmt::<unnamed>::gemm_bf16_tcgen05_tma_kernel<(int)0, (int)2>(CUtensorMap_st)", 10.0, 100, 50)
    add(2, "void fmmt::<unnamed>::gemm_bf16_tcgen05_tma_kernel<(int)0, (int)2>(CUtensorMap_st)", 30.0, 300, 150)
    add(3, "void fmmt::<unnamed>::layernorm_vec_kernel<(int)32, (int)6>(fmmt::LnParams)", 60.0, 10, 10)
    csv_path.write_text("\n".join(rows) + "\n")
    out = tmp_path / "out.json"
    tool.launches(str(csv_path), str(out), 1)
    d = json.load(open(out))
    assert d["launches"] == 3 and abs(d["total_us"] - 100.0) < 1e-6
    g = next(k for k in d["kernels"] if k["kernel"].startswith("gemm_bf16_tcgen05_tma_kernel"))
    assert g["launches"] == 2 and abs(g["us"] - 40.0) < 1e-6 and abs(g["share"] - 0.4) < 1e-6
    assert abs(g["dram_MB_per_launch"] - 300.0) < 1e-6                          # (150 + 450) MB over two launches
