"""7x7/s2 stem conv as a 5x5/s1 conv on the pixel-unshuffled input (Cin*4 channels): timing + exactness."""
import torch, torch.nn.functional as F
dev = torch.device('cuda:0')
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n * 1e3
def s2d_weight(w7):
    co, ci = w7.shape[:2]
    w5 = torch.zeros(co, ci * 4, 5, 5, dtype=w7.dtype, device=w7.device)
    for ky in range(7):
        u = ky - 3; a = u // 2; p = u - 2 * a
        for kx in range(7):
            v = kx - 3; b = v // 2; q = v - 2 * b
            w5[:, p * 2 + q::4, a + 2, b + 2] = w7[:, :, ky, kx]
    return w5
torch.manual_seed(0)
for (n, cin) in ((5, 5), (1, 3)):
    x = torch.randn(n, cin, 480, 864, device=dev)
    w7 = torch.randn(64, cin, 7, 7, device=dev) * 0.05
    bias = torch.randn(64, device=dev)
    torch.backends.cudnn.allow_tf32 = False
    want = F.relu(F.conv2d(x, w7, bias, 2, 3))
    x2 = F.pixel_unshuffle(x, 2).contiguous(memory_format=torch.channels_last)
    w5 = s2d_weight(w7).contiguous(memory_format=torch.channels_last)
    got = torch.cudnn_convolution_relu(x2, w5, bias, (1, 1), (2, 2), (1, 1), 1)
    print('n', n, 'cin', cin, 'max abs err fp32', (got - want).abs().max().item(), 'rel', ((got - want).abs().max() / want.abs().max()).item())
    torch.backends.cudnn.allow_tf32 = True
    xcl = x.contiguous(memory_format=torch.channels_last); w7cl = w7.contiguous(memory_format=torch.channels_last)
    print('   7x7/s2 direct :', t(lambda: torch.cudnn_convolution_relu(xcl, w7cl, bias, (2, 2), (3, 3), (1, 1), 1)), 'us')
    print('   5x5 on s2d    :', t(lambda: torch.cudnn_convolution_relu(x2, w5, bias, (1, 1), (2, 2), (1, 1), 1)), 'us')
    print('   pixel_unshuffle + channels_last copy:', t(lambda: F.pixel_unshuffle(x, 2).contiguous(memory_format=torch.channels_last)), 'us')
    # pad input channels to a multiple of 8
    c2 = x2.shape[1]; c8 = (c2 + 7) // 8 * 8
    if c8 != c2:
        x2p = F.pad(x2, (0, 0, 0, 0, 0, c8 - c2)).contiguous(memory_format=torch.channels_last)
        w5p = F.pad(w5, (0, 0, 0, 0, 0, c8 - c2)).contiguous(memory_format=torch.channels_last)
        print(f'   5x5 on s2d, Cin padded {c2}->{c8}:', t(lambda: torch.cudnn_convolution_relu(x2p, w5p, bias, (1, 1), (2, 2), (1, 1), 1)), 'us')
