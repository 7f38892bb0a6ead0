"""Per-conv roofline table of the ResNet-50 trunk (model_copenet.py:161-176) for n images.
    python tools/trunk_roofline.py [n_images] [tflops] [hbm_gbs]
Algorithmic bytes = input read once + output written once + weights (bf16) + residual read."""
import sys

def specs():
    v = [("stem", 64, 3, 7, 2, 3, 224)]
    layers, planes = [3, 4, 6, 3], [64, 128, 256, 512]
    inpl, H = 64, 56
    for li in range(4):
        for b in range(layers[li]):
            s = 2 if (li > 0 and b == 0) else 1
            p = "layer%d.%d" % (li + 1, b)
            v.append((p + ".conv1", planes[li], inpl, 1, 1, 0, H))
            v.append((p + ".conv2", planes[li], planes[li], 3, s, 1, H))
            if b == 0:
                v.append((p + ".down", planes[li] * 4, inpl, 1, s, 0, H))
            H //= s
            v.append((p + ".conv3", planes[li] * 4, planes[li], 1, 1, 0, H))
            inpl = planes[li] * 4
    return v

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    tf = float(sys.argv[2]) if len(sys.argv) > 2 else 1384.5
    bw = float(sys.argv[3]) if len(sys.argv) > 3 else 6552.3
    tot_f = tot_b = tot_t = 0.0
    print("%-18s %8s %5s %5s %9s %8s %8s %8s %6s" % ("conv", "M", "N", "K", "GFLOP", "MB", "us_tc", "us_hbm", "tiles"))
    for name, co, ci, k, s, p, H in specs():
        Ho = (H + 2 * p - k) // s + 1
        M, N, K = n * Ho * Ho, co, ci * k * k
        fl = 2.0 * M * N * K
        by = n * H * H * ci * (4 if name == "stem" else 2) + M * N * 2 + N * K * 2
        if name.endswith("conv3"):
            by += M * N * 2
        t_tc, t_hbm = fl / tf / 1e6, by / bw / 1e3
        tiles = -(-M // 128) * -(-N // 128)
        print("%-18s %8d %5d %5d %9.3f %8.2f %8.2f %8.2f %6d" % (name, M, N, K, fl / 1e9, by / 1e6, t_tc, t_hbm, tiles))
        tot_f += fl; tot_b += by; tot_t += max(t_tc, t_hbm)
    print("total %.2f GFLOP (%.3f/img)  %.1f MB (%.2f/img)  tensor-only %.1f us, hbm-only %.1f us, sum of per-layer max %.1f us"
          % (tot_f / 1e9, tot_f / 1e9 / n, tot_b / 1e6, tot_b / 1e6 / n, tot_f / tf / 1e6, tot_b / bw / 1e3, tot_t))

if __name__ == "__main__":
    main()
