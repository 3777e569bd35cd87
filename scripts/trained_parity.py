"""Run the trained-weights parity protocol (tests/trained_protocol.py) and write the report.
usage: python scripts/trained_parity.py [steps] [n_videos] [out.json] [emulate: tf32,fp16,fp16-body,bf16|none] [lr] [strength]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tests import trained_protocol as TP  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
n_videos = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
out_path = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/r2_trained_parity.json"
emu = tuple(m for m in (sys.argv[4] if len(sys.argv) > 4 else "tf32,fp16").split(",") if m and m != "none")
lr = float(sys.argv[5]) if len(sys.argv) > 5 else 2e-4
strength = float(sys.argv[6]) if len(sys.argv) > 6 else 1.0
dev = torch.device("cuda:0")
t0 = time.time()
eng, tr, protos, losses = TP.train_model(dev, steps=steps, lr=lr, strength=strength)
t1 = time.time()
rep = TP.evaluate(eng, protos, dev, n_videos=n_videos, emulate=emu, strength=strength, fp64=True)
rep.update(train_steps=steps, lr=lr, strength=strength, losses=losses, train_seconds=t1 - t0, eval_seconds=time.time() - t1,
           skipped_steps=int(tr.skipped_steps()) if hasattr(tr, "skipped_steps") else None,
           shape=TP.SHAPE, host_cores=os.cpu_count())
os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
json.dump(rep, open(out_path, "w"), indent=1)
print(json.dumps(rep, indent=1))
