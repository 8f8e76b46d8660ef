"""GPU diagnostic: engine vs oracle on small configs — loss, metrics, gradients per tensor, train step, sampling."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lap_b200.config import get_config
from lap_b200 import params as P, ops
from lap_b200.data import synthetic_batch
from lap_b200.model import LAP
from lap_b200.observation import CoTObservation, Observation
from lap_b200.train import init_train_state, TrainingStepRunner, batch_from_dict
from oracle import lap_oracle as O

name = sys.argv[1] if len(sys.argv) > 1 else "debug_tiny"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
tc = get_config(name); cfg = tc.model
ref = P.init_reference_params(cfg, 0, reference_zero_init=False)
b = synthetic_batch(cfg, B, step=1)
t = lambda x: torch.from_numpy(np.asarray(x))
obs_o = dict(images={k: t(v) for k, v in b["image"].items()}, image_masks={k: t(v) for k, v in b["image_mask"].items()},
             tokenized_prompt=t(b["tokenized_prompt"]), tokenized_prompt_mask=t(b["tokenized_prompt_mask"]),
             tokenized_langact_mask=t(b["tokenized_langact_mask"]), token_loss_mask=t(b["token_loss_mask"]), sample_mask=t(b["sample_mask"]))
acts, noise, tm = t(b["actions"]), t(b["noise"]), t(b["time"])

def rel(a, b_):
    a = a.float().cpu(); b_ = b_.float().cpu()
    return ((a - b_).norm() / b_.norm().clamp_min(1e-30)).item()

model = LAP(cfg, init=False); model.load_params(ref)
obs, actions, extra = batch_from_dict(b)
# ---- masks bit-exact ----
st = model._stage(obs, actions, noise, tm, with_loss=True)
loss_e, m_e = model.compute_loss(0, obs, actions, noise=noise, time=tm)
Pn, A = cfg.prefix_len, cfg.action_horizon; T = Pn + A; Tpad = (T + 63) // 64 * 64
bits = model._bufs["mask.bits"]; dense = torch.zeros(B, T, T, dtype=torch.uint8, device="cuda")
ops.mask_expand(bits, dense, B * T, T, Tpad // 32)
pos_e = model._bufs["mask.pos"].cpu()
for bf in (True, False):
    loss_o, m_o, aux = O.compute_loss(ref, cfg, obs_o, acts, noise, tm, bf16=bf, return_aux=True)
    print(f"[{name}] oracle bf16={bf}: loss {loss_o.item():.6f} engine {loss_e.item():.6f} rel {abs(loss_o.item()-loss_e.item())/abs(loss_o.item()):.2e}")
    for k in m_o: print(f"    {k}: oracle {m_o[k].item():.6f} engine {m_e[k].item():.6f}")
print("mask equal:", torch.equal(dense.cpu().bool(), aux["mask"]), " positions equal:", torch.equal(pos_e, aux["positions"]))
# intermediate: prefix tokens (siglip + embed)
X0 = model._bufs["g.X.tmp0"].float().cpu().view(B, Pn, -1)
loss_o, m_o, auxb = O.compute_loss(ref, cfg, obs_o, acts, noise, tm, bf16=True, return_aux=True)
print("prefix tokens rel (vs bf16 oracle):", rel(X0, auxb["prefix_tokens"]), " image part:", rel(X0[:, :Pn - cfg.max_token_len], auxb["prefix_tokens"][:, :Pn - cfg.max_token_len]))
print("suffix tokens rel:", rel(model._bufs["g.XE.tmp0"].view(B, A, -1), auxb["suffix_tokens"]), " cond rel:", rel(model._bufs["suf.cond"], auxb["cond"]))
print("v_t rel:", rel(model._bufs["loss.v"].view(B, A, -1), auxb["v_t"]))
# ---- gradients ----
state = init_train_state(tc, model=model)
runner = TrainingStepRunner(tc)
st = model._stage(obs, actions, noise, tm, with_loss=True)
loss = model.forward_backward(st)
torch.cuda.synchronize()
g_eng = model.params_reference(model.G)
for bf in (True, False):
    ostate = dict(step=0, params=ref, mu={k: torch.zeros_like(v) for k, v in ref.items()}, nu={k: torch.zeros_like(v) for k, v in ref.items()}, ema={k: v.clone() for k, v in ref.items()})
    ns, info_o, g_o = O.train_step(tc, ostate, obs_o, acts, noise, tm, bf16=bf)
    tot_e = torch.sqrt(sum((v.double() ** 2).sum() for v in g_eng.values())).item()
    print(f"grad check vs oracle bf16={bf}: |g| oracle {info_o['grad_norm'].item():.5f} engine {tot_e:.5f}")
    worst = []
    for k in g_o:
        r = rel(g_eng[k], g_o[k]); worst.append((r, k, g_o[k].norm().item()))
    worst.sort(reverse=True)
    for r, k, nrm in worst[:60 if bf else 12]: print(f"    {r:.3e}  |g|={nrm:.3e}  {k}")
# ---- full train step ----
model.load_params(ref)
state = init_train_state(tc, model=model)
state, info = runner(0, state, (obs, actions, extra))
torch.cuda.synchronize()
ostate = dict(step=0, params=ref, mu={k: torch.zeros_like(v) for k, v in ref.items()}, nu={k: torch.zeros_like(v) for k, v in ref.items()}, ema={k: v.clone() for k, v in ref.items()})
ns, info_o, _ = O.train_step(tc, ostate, obs_o, acts, noise, tm, bf16=True)
print("train step info engine:", {k: round(float(v), 6) for k, v in info.items()})
print("train step info oracle:", {k: round(float(v), 6) for k, v in info_o.items()})
p_eng = model.params_reference()
dp = [(rel(p_eng[k] - ref[k], ns["params"][k] - ref[k]), k) for k in ref]
dp.sort(reverse=True); print("param update rel err (worst 8):", [(f"{r:.2e}", k.split('/')[-3:]) for r, k in dp[:8]])
ema_eng = model.params_reference(state.ema_params)
print("ema rel:", max(rel(ema_eng[k], ns["ema"][k]) for k in ref))
# ---- sampling ----
model.load_params(ref)
b2 = dict(b); b2.pop("tokenized_langact_mask")
obs_inf = Observation.from_dict(b2)
t0 = time.time(); a_e = model.sample_actions(0, obs_inf, num_steps=10, noise=noise); torch.cuda.synchronize(); t1 = time.time()
obs_o2 = dict(obs_o); obs_o2["tokenized_langact_mask"] = None
for bf in (True, False):
    a_o = O.sample_actions(ref, cfg, obs_o2, noise, num_steps=10, bf16=bf)
    print(f"sample_actions rel vs oracle bf16={bf}: {rel(a_e, a_o):.3e}  max abs {((a_e.cpu()-a_o).abs().max()).item():.3e}")
print("launches so far:", ops.launch_count, " sample time", t1 - t0)
print("DONE")
