"""Development aid (run under gpurun): the rigid-body front end alone, configs[2] shape, 65 536 states."""
import sys, torch
sys.path.insert(0, ".")
from qppvm_b200 import api, gen
from qppvm_b200.layout import CONFIGS
ci = int(sys.argv[1]) if len(sys.argv) > 1 else 2
d = CONFIGS[ci]["desc"]; rob = gen.robot_for(d.n_a); s = api.Solver(d); s.set_robot(rob, (rob.foot + rob.hand)[:d.n_contacts])
st = torch.from_numpy(gen.generate_states(d, 4096, 1)).cuda().repeat(16, 1).contiguous()
rec = s.records_from_states(st)
for _ in range(3):
    s.records_from_states(st, records=rec)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    s.records_from_states(st, records=rec)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("rbd_records_kernel config %d: %d states in %.3f ms = %.1f M states/s, %.0f GB/s of record bytes" % (ci, st.shape[0], ms, st.shape[0] / ms / 1e3, rec.numel() * 8 / ms / 1e6))
