"""Small-shape debug of the streaming gradient kernel (prints progress; run under `timeout`)."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("camera-aware-neural-networks-for-few-view-depth-estimation_b200")
dev = torch.device("cuda:0")
for (B, H, W) in [(1, 16, 128), (2, 48, 136), (4, 96, 256), (32, 480, 640)]:
    b = pkg.synth.make_batch(B, H, W, seed=7, device=dev)
    for terms in (pkg.TERM_GRAD, 7, 15):
        params = pkg.default_params(terms=terms)
        outs = []
        for mode in (128, 0):
            pkg.force_generic(mode)
            print(f"shape {(B, H, W)} terms {terms} mode {mode} ...", flush=True)
            ws = pkg.stack_fwd_bwd(b["pred"], b["gt"], b["rgb"], b["K"], None, params=params)
            torch.cuda.synchronize()
            rr = ws.read_results()
            print('   dbg', [int.from_bytes(bytes(memoryview(rr._pad0))[4*i:4*i+4], 'little', signed=True) for i in range(3)], flush=True)
            outs.append((ws.grad.clone(), pkg.results_dict(rr)))
        pkg.force_generic(0)
        d = (outs[0][0] - outs[1][0]).abs().max().item() / max(outs[0][0].abs().max().item(), 1e-30)
        print(f"   grad diff {d:.3e}  loss {outs[0][1]['loss_total']:.8f} / {outs[1][1]['loss_total']:.8f}", flush=True)
