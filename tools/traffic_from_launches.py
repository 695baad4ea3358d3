#!/usr/bin/env python
"""DRAM traffic of the dominant kernel from an ncu pass that collected
gpu__time_duration.sum, dram__bytes_read.sum and dram__bytes_write.sum for one training step:
average bytes per launch over the 63 ConvLayer forward launches (the k_conv_tc launches between k_stem_fwd and
k_loss_sqerr).  Writes profiles/ncu_traffic.json, which bench.py quotes as roofline.traffic.
usage: python tools/traffic_from_launches.py launches_dram.csv [step] > profiles/ncu_traffic.json"""
import json
import sys

sys.path.insert(0, __file__.rsplit('/', 1)[0])
from launch_summary import load  # noqa: E402


def unit_scale(path):
    return 1.0


def main():
    path = sys.argv[1]
    step = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    dur = load(path)
    rd = load(path, 'dram__bytes_read.sum')
    wr = load(path, 'dram__bytes_write.sum')
    assert len(dur) == len(rd) == len(wr), (len(dur), len(rd), len(wr))
    stems = [i for i, l in enumerate(dur) if l[0].startswith('k_stem_fwd')]
    i0 = stems[step]
    i1 = next(i for i in range(i0, len(dur)) if dur[i][0].startswith('k_loss_sqerr'))
    idx = [i for i in range(i0, i1) if dur[i][0].startswith('k_conv_tc') or dur[i][0].startswith('k_igemm')]
    tot_r = sum(rd[i][1] for i in idx)
    tot_w = sum(wr[i][1] for i in idx)
    tot_t = sum(dur[i][1] for i in idx)
    print(json.dumps({
        "source": path, "step": step, "launches": len(idx),
        "conv_fwd_avg_bytes_per_launch": (tot_r + tot_w) / len(idx),
        "conv_fwd_dram_read_bytes": tot_r, "conv_fwd_dram_write_bytes": tot_w,
        "conv_fwd_avg_ns_per_launch_under_ncu": tot_t / len(idx),
        "note": "dram__bytes_read.sum + dram__bytes_write.sum summed over the 63 forward ConvLayer launches of one "
                "training step, divided by 63; ncu serialises the kernels, so each launch starts with the previous "
                "layer's output still in the 126 MB L2 exactly as in the real step"}))


if __name__ == '__main__':
    main()
