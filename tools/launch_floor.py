import torch, numpy as np
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
one = torch.zeros(1, device="cuda")
big = torch.zeros(1<<20, device="cuda")
def t(fn, fl=True, reps=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    ts=[]
    for _ in range(reps):
        if fl: flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1)*1e3)
    return float(np.median(ts))
print("empty, after flush", t(lambda: None))
print("empty, no flush", t(lambda: None, False))
print("1-elem add, after flush", t(lambda: one.add_(1)))
print("1-elem add, no flush", t(lambda: one.add_(1), False))
print("1M add, after flush", t(lambda: big.add_(1)))
