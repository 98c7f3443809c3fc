"""Host-side per-frame time of the lookahead session for 1..16 concurrent streams."""
import os, sys, threading, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from x264vfw_b200 import lookahead
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from clipgen import SyntheticClip

W, H, N = 1920, 1080, 24
clip = SyntheticClip(W, H, n_frames=N, cuts=(15,), flash=None)
HOST = int(os.environ.get("HOST", "0"))      # 1: pinned host frames in, converted planes out (the e2e path)
if HOST:
    frames = [torch.from_numpy(clip.packed(i, "bgra")).pin_memory() for i in range(N)]
    dev_frames = [f.cuda() for f in frames]
else:
    frames = [torch.from_numpy(clip.packed(i, "bgra")).cuda() for i in range(N)]
torch.cuda.synchronize()

BG = int(os.environ.get("BGCOPY", "0"))     # 1: a background thread keeps the H2D link busy (diagnostics)
if BG:
    bg_h = torch.empty(8294400, dtype=torch.uint8).pin_memory(); bg_d = torch.empty(8294400, dtype=torch.uint8, device="cuda")
    bg_stop = False
    def bg():
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            while not bg_stop:
                for _ in range(BG):
                    bg_d.copy_(bg_h, non_blocking=True)
                st.synchronize()
                time.sleep(0.0002 if BG < 8 else 0)
    threading.Thread(target=bg, daemon=True).start()

def run(S, total=int(os.environ.get("TOTAL", "120"))):
    las = [lookahead.Lookahead(lookahead.params_preset("medium", W, H), in_csp=9 | 0x1000, device=0) for _ in range(S)]
    base = {}
    def work(la):
        conv = [torch.empty(W * H * 3 // 2, dtype=torch.uint8).pin_memory().numpy() for _ in range(2)] if HOST else None
        for i in range(total):
            if i == 60:
                base[id(la)] = (la.counters(), time.perf_counter())
            if HOST == 2:      # H2D only
                la.put_frame(frames[i % N].numpy(), on_device=False)
            elif HOST == 3:    # D2H only
                la.put_frame(dev_frames[i % N].data_ptr(), on_device=True, conv_pic=conv[i & 1])
            elif HOST:
                la.put_frame(frames[i % N].numpy(), on_device=False, conv_pic=conv[i & 1])
            else:
                la.put_frame(frames[i % N].data_ptr(), on_device=True)
            la.decisions()
    ths = [threading.Thread(target=work, args=(la,)) for la in las]
    t0 = time.perf_counter()
    for t in ths: t.start()
    for t in ths: t.join()
    dt = time.perf_counter() - t0
    c1 = las[0].counters(); c0, tb = base[id(las[0])]
    c = {k: c1[k] - c0[k] for k in c1}
    fps = S * c['frames'] / (time.perf_counter() - tb)
    print(f"S={S}: {S*total/dt:8.1f} fps overall, steady {fps:8.1f} fps  per-frame host us: put {c['put_us']/c['frames']:.0f} decide {c['decide_us']/c['frames']:.0f} sync {c['sync_us']/c['frames']:.0f}  launches/frame {c['launches']/c['frames']:.1f} syncs/frame {c['syncs']/c['frames']:.2f}")
    for la in las: la.close()

for S in [int(x) for x in os.environ.get("STREAMS", "1,2,4,8,16").split(",")]:
    run(S)
