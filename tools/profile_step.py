"""One eager (un-graphed) training step at a given batch, for ncu: every kernel of the step is launched exactly once per step.

  ncu --set full --import-source on --clock-control none -k regex:slide_conv_kernel --launch-skip 16 --launch-count 1 \
      -o gpurun_out/prof_x python tools/profile_step.py --batch 1024 --steps 1
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=1024)
    ap.add_argument('--steps', type=int, default=1)
    a = ap.parse_args()
    import torch
    import wiflow_b200 as wf
    from oracle import wiflow_oracle as O          # synthetic inputs only
    dev = torch.device('cuda', 0)
    torch.manual_seed(0)
    model = wf.WiFlowPoseModel(dropout=0.5).to(dev)
    ts = wf.TrainStep(model, a.batch, use_cuda_graph=False)
    x, y = O.synthetic_batch(a.batch, seed=1)
    x, y = x.to(dev), y.to(dev)
    for _ in range(a.steps):
        out = ts.step(x, y)
    torch.cuda.synchronize(dev)
    print('loss', out.tolist())


if __name__ == '__main__':
    main()
