import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
from conftest import get_native_tts
from oracle import tts_oracle as T
m = get_native_tts(); dev = m.device; nat = m.native
sd = T.synthetic_tts_state_dict()
for B, Tn, lens in ((1, 300, [300]), (2, 600, [600, 257]), (2, 1500, [1500, 700])):
    tokens, lengths, sid, noise_w = T.synthetic_tts_inputs(B, Tn, 21, lens)
    nat.debug_enable(True)
    yl, w_ceil, logw = nat.tts_encode(tokens.to(dev), lengths.to(dev), sid.to(dev), noise_w=noise_w.to(dev), noise_scale_w=0.6, length_scale=1.0, sdp_ratio=0.2)
    torch.cuda.synchronize()
    x = nat.debug_fetch("tts.x"); l0 = nat.debug_fetch("tts.layer0"); st = nat.debug_fetch("tts.stats"); xc = nat.debug_fetch("tts.sdp_cond")
    ls, ld = nat.debug_fetch("tts.logw_sdp")[:, 0], nat.debug_fetch("tts.logw_dp")[:, 0]
    nat.debug_enable(False)
    with torch.no_grad():
        rx, rm, rl, mask = T.text_encoder(sd, tokens, lengths)
        g = sd["emb_g.weight"][sid].unsqueeze(-1)
        rs = T.sdp_reverse(sd, rx, mask, g, noise_w, 0.6); rd = T.duration_predictor(sd, rx, mask, g)
    mk = mask[:, 0].numpy()
    ex = np.abs(x * mk[:, :, None] - rx.transpose(1, 2).numpy()).max(axis=2)
    es = np.abs(ls * mk - rs[:, 0].numpy() * mk); ed = np.abs(ld * mk - rd[:, 0].numpy() * mk)
    print("T", Tn, "x max", ex.max(), "argmax", np.unravel_index(ex.argmax(), ex.shape), "dp max", ed.max(), np.unravel_index(ed.argmax(), ed.shape),
          "sdp max", es.max(), np.unravel_index(es.argmax(), es.shape), "n>5e-4", int((es > 5e-4).sum()), "n>1e-2", int((es > 1e-2).sum()))
    top = np.argsort(es.ravel())[-8:]
    print("  top sdp err idx", [(int(i // Tn), int(i % Tn), float(es.ravel()[i])) for i in top])
    print("  x err by 256-block", [float(ex[0, i:i + 256].max()) for i in range(0, Tn, 256)])
