"""Synthetic REDD-like load profiles and friends.

Integer centi-watts k are generated first, then x = float32(k) / float32(100) so that the reference's
normalize (x * 100 + 0.5, truncated) recovers k.  Runs on any torch device; seeded per channel chunk.
"""
from __future__ import annotations

import torch


def redd_like(n_channels: int, n_samples: int, device="cpu", seed: int = 0xDC0000, chunk: int = 128,
              appliances: int = 8, return_int: bool = False) -> torch.Tensor:
    """float32 [n_channels, n_samples]: base load U[50,400] W + 8 two-state Markov appliances
    (p_on in [1/7200,1/600], p_off in [1/3600,1/300] per second, power U[20,2500] W) + round(N(0,0.5 W)) noise,
    in centi-watts, clipped to [0, 2^24)."""
    device = torch.device(device)
    out = torch.empty((n_channels, n_samples), dtype=torch.int32 if return_int else torch.float32, device=device)
    n_events = 1024
    for c0 in range(0, n_channels, chunk):
        c = min(chunk, n_channels - c0)
        g = torch.Generator(device=device)
        g.manual_seed(seed + c0)

        def U(shape, lo, hi):
            return lo + (hi - lo) * torch.rand(shape, generator=g, device=device, dtype=torch.float64)

        base = torch.round(100.0 * U((c, 1), 50.0, 400.0)).to(torch.int64)
        p_on = U((c, appliances, 1), 1.0 / 7200, 1.0 / 600)
        p_off = U((c, appliances, 1), 1.0 / 3600, 1.0 / 300)
        power = torch.round(100.0 * U((c, appliances), 20.0, 2500.0)).to(torch.int64)
        # alternating geometric sojourn times, starting "off"
        u = torch.rand((c, appliances, n_events), generator=g, device=device, dtype=torch.float64).clamp_min(1e-300)
        p = torch.where((torch.arange(n_events, device=device) % 2 == 0).view(1, 1, -1), p_on, p_off)
        dur = torch.floor(torch.log(u) / torch.log1p(-p)).to(torch.int64) + 1
        t_event = torch.cumsum(dur, dim=2)
        toggles = torch.zeros((c, appliances, n_samples + 1), dtype=torch.int32, device=device)
        toggles.scatter_add_(2, t_event.clamp_max(n_samples), torch.ones_like(t_event, dtype=torch.int32))
        state = torch.cumsum(toggles[:, :, :n_samples], dim=2) & 1
        k = base + (state.to(torch.int64) * power.unsqueeze(2)).sum(dim=1)
        noise = torch.round(50.0 * torch.randn((c, n_samples), generator=g, device=device, dtype=torch.float32)).to(torch.int64)
        k = (k + noise).clamp_(0, (1 << 24) - 1)
        if return_int:
            out[c0:c0 + c] = k.to(torch.int32)
        else:
            out[c0:c0 + c] = k.to(torch.float32) / torch.tensor(100.0, dtype=torch.float32, device=device)
    return out


def meters_15min(n_meters: int, n_samples: int = 96, device="cpu", seed: int = 0xA0D2, chunk: int = 65536) -> torch.Tensor:
    """float32 [n_meters, n_samples]: 15-minute energy in Wh with two decimals:
    daily double-hump profile x per-meter LogNormal(0, 0.5) scale, k in [0, 2e5)."""
    device = torch.device(device)
    out = torch.empty((n_meters, n_samples), dtype=torch.float32, device=device)
    t = torch.arange(n_samples, device=device, dtype=torch.float32) / n_samples * 24.0
    hump = 0.25 + torch.exp(-0.5 * ((t - 7.5) / 1.5) ** 2) + 1.4 * torch.exp(-0.5 * ((t - 19.0) / 2.2) ** 2)
    for m0 in range(0, n_meters, chunk):
        m = min(chunk, n_meters - m0)
        g = torch.Generator(device=device)
        g.manual_seed(seed + m0)
        scale = torch.exp(0.5 * torch.randn((m, 1), generator=g, device=device))
        jitter = 1.0 + 0.15 * torch.randn((m, n_samples), generator=g, device=device)
        wh = (150.0 * scale * hump.view(1, -1) * jitter).clamp_(0.0, 1999.99)
        k = torch.round(wh * 100.0).to(torch.int64).clamp_(0, 199999)
        out[m0:m0 + m] = k.to(torch.float32) / torch.tensor(100.0, dtype=torch.float32, device=device)
    return out


def watts_u16(n_channels: int, n_samples: int, device="cpu", seed: int = 0xDC0000) -> torch.Tensor:
    """uint8 [n_channels, 2 * n_samples]: integer-watt profiles as little-endian uint16 (config 4, LZMH input)."""
    k = redd_like(n_channels, n_samples, device=device, seed=seed, return_int=True)
    w = ((k + 50) // 100).clamp_(0, 65535).to(torch.int32)
    lo = (w & 0xFF).to(torch.uint8)
    hi = (w >> 8).to(torch.uint8)
    return torch.stack((lo, hi), dim=2).reshape(n_channels, 2 * n_samples).contiguous()


def lzmh_blocks(n_blocks: int, block_bytes: int = 65536, device="cpu", seed: int = 0xDC0000) -> torch.Tensor:
    """uint8 [n_blocks * block_bytes]: the watts_u16 profiles cut into LZMH blocks (config 4), with one adjustment the
    survey asks for: no block ends in a zero byte.  The reference's decoder drops trailing zero-valued literals whose bits
    sit in its 32-bit window when the input ends (lzmh.c:571; `"A\0"` decodes to `"A"`), so such a block does not
    round-trip through the REFERENCE either; a last high byte of 0 (a reading below 256 W) becomes 1."""
    per_channel = 2 * 86400
    chans = (n_blocks * block_bytes + per_channel - 1) // per_channel
    data = watts_u16(chans, 86400, device=device, seed=seed).reshape(-1)[: n_blocks * block_bytes].contiguous()
    data.view(n_blocks, block_bytes)[:, -1].clamp_(min=1)
    return data
