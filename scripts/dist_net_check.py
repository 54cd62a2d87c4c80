"""torchrun script: sharded encrypted inference over NCCL must reproduce the single-GPU ciphertexts bit for bit.
usage: torchrun --nproc-per-node N scripts/dist_net_check.py [net]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import redsec_b200 as rs
from redsec_b200 import client, netspec, nets

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
name = sys.argv[1] if len(sys.argv) > 1 else "mnist/sign1024x1"
spec = netspec.NETS[name]()
ks = client.keygen(0)
eng = rs.Engine(local); eng.load_eval_key(ks.bsk, ks.ksk)
label, px = netspec.load_image_csv(spec["image"])
ct = client.encrypt(netspec.map_pixels(spec, px) * client.UNIT, ks.lwe_key, client.SECALPHA, 7)
net = nets.EncryptedNet(eng, spec)
net.build_tables(rank, world)
d = eng.upload(ct)
out = net.run(d, dist=(dist, rank, world)); eng.sync()
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
out = net.run(d, dist=(dist, rank, world)); eng.sync()
dist.barrier(); torch.cuda.synchronize()
dt = time.perf_counter() - t0
sharded = eng.download(out)
scores = client.decrypt(sharded, ks.lwe_key, 4096)
same = None
if name.startswith("mnist") or os.environ.get("RS_CHECK_SINGLE"):
    single = eng.download(net.run(d))
    same = bool(np.array_equal(single, sharded))
if os.environ.get("RS_LAYER_TIMES"):
    times = []
    net.run(d, dist=(dist, rank, world), times=times).free()
    if rank == 0:
        print("layer seconds (rank 0):", [round(t, 4) for t in times], "sum", round(sum(times), 4))
if rank == 0:
    print({"net": name, "world": world, "s_per_image": dt, "bootstraps": net.bootstraps(), "argmax": int(np.argmax(scores)),
           "label": label, "scores": scores.tolist(), "sharded_equals_single_gpu": same})
dist.destroy_process_group()
