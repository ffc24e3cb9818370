dev = torch.device("cuda", 0)
B = int(os.environ.get("B", 16))
torch.manual_seed(0)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)) * 1e3


def planes(f32):
    p = ops.Planes(BF16X2, f32.shape, dev)
    p.p0.copy_(f32.to(torch.bfloat16))
    p.p1.copy_((f32 - p.p0.float()).to(torch.bfloat16))
    return p


x = planes(torch.randn(B, 320, 320, 64, device=dev))
r = planes(torch.randn(B, 320, 320, 64, device=dev))
w = torch.randn(64, 64, 3, 3, device=dev) * 0.05
_, whi, wlo = ops.pack_conv_weight(w, simt=False, tc=True, split=True)
sc, sh = torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev)
NSM = ops.device_info()[0]
dbg = torch.zeros(8 * NSM, dtype=torch.int64, device=dev)


def conv(res, out_dtype=None):
    return ops.conv3x3_bn_act_fwd(x, whi, wlo, sc, sh, res=res, relu=True, out_dtype=out_dtype, engine=TC)


def breakdown(res):
    dbg.zero_()
    ops.debug_buffer(dbg)
    conv(res)
    torch.cuda.synchronize()
    ops.debug_buffer(None)
    d = dbg.view(NSM, 8).double().cpu().numpy()
    tot = d[:, 4].mean()
    names = ["issuer waits operands", "issuer waits accumulator", "producer waits slot", "epilogue(w2) waits accumulator"]
    return "  ".join(f"{n} {d[:, i].mean() / tot:5.1%}" for i, n in enumerate(names)) + f"  | CTA cycles {tot:,.0f}, tiles/CTA {d[:, 5].mean():.1f}"


