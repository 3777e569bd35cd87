import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learnablepoolingmethods_b200 import variables, ops
from learnablepoolingmethods_b200.engine import NetVladConfig, NetVladEngine
from oracle import netvlad_oracle as O
from tests.helpers import oracle_params, perturb, rel
dev = torch.device("cuda:0")
B, K, Hd, V, T, D = 4, 64, 64, 100, 256, 1024
store = variables.VariableStore(dev, seed=1810)
eng = NetVladEngine(NetVladConfig(model="NetVladV2", iterations=T, cluster_size=K, hidden_size=Hd, vocab_size=V), store)
perturb(store)
P, S = oracle_params(store)
x, nf, _ = O.synthetic_batch(B, seed=20181000, vocab=V)
v, sh = store.vars, eng.refresh_shadows()
# oracle stages (inference)
with torch.no_grad():
    xs = O.sample_uniform_frames(x, nf, T).reshape(-1, 1152)
    xb = O.batch_norm(xs, P, S, "input_bn", False)[:, :D]
    X3 = xb.reshape(B, T, D)
    a = "video_VLAD/cluster_attention"; H = D // 16
    q = O.dense(X3, P, a + "/q", bias=False); k = O.dense(X3, P, a + "/k", bias=False); vv = O.dense(X3, P, a + "/v", bias=False)
    qh, kh, vh = O._split_heads(q, H), O._split_heads(k, H), O._split_heads(vv, H)
    logits = qh @ kh.transpose(-1, -2)
    lb = O.batch_norm(logits, P, S, a + "/logits_bn", False)
    o_ref = O._combine_heads(torch.softmax(lb, -1) @ vh)
    obn_ref = O.batch_norm(o_ref, P, S, a + "/attention_bn", False)
    att_ref = O.dense(obn_ref, P, a + "/output_transform")
    h1_ref = O.layer_norm_joint(att_ref + X3, P, a + "/LayerNorm")
    f_ref = O.dense(h1_ref, P, a + "/filter_outputencode", relu=True)
    fbn_ref = O.batch_norm(f_ref, P, S, a + "/filter_bn", False)
    f2_ref = O.dense(fbn_ref, P, a + "/ff_outputencode", relu=True)
    A_ref = O.batch_norm(f2_ref, P, S, a + "/feed_output_bn", False)
    C = P["video_VLAD/cluster_centers"]
    Vun = torch.matmul(X3.transpose(1, 2), A_ref) - A_ref.sum(1, keepdim=True) * C      # [B,D,K]
# engine stages
r = ops.bn_finalize(None, None, 1, v["input_bn/gamma"], v["input_bn/beta"], v["input_bn/moving_mean"], v["input_bn/moving_variance"], training=False, bessel=True)
Xr, Xa = ops.sample_bn_apply(x.to(dev), nf.to(dev), T, r[0], r[1], split_col=1024)
print("X", rel(Xr.float(), xb))
qkv = ops.gemm(Xr, sh[a + "/wqkv16"]); print("qkv", rel(qkv.float(), torch.cat([q, k, vv], -1).reshape(B * T, -1)))
bn = a + "/logits_bn"
ks, kb = ops.bn_finalize(None, None, 1, v[bn + "/gamma"], v[bn + "/beta"], v[bn + "/moving_mean"], v[bn + "/moving_variance"], training=False, bessel=True)
o = ops.mha_core_fwd(qkv, B, T, D, H, scale=1.0, key_scale=ks, key_shift=kb); print("o", rel(o.float(), o_ref.reshape(B * T, D)))
bn = a + "/attention_bn"
ops.batch_norm_cols_f16(o, v[bn + "/gamma"], v[bn + "/beta"], v[bn + "/moving_mean"], v[bn + "/moving_variance"], training=False, bessel=False); print("o_bn", rel(o.float(), obn_ref.reshape(B * T, D)))
att = ops.gemm(o, sh[a + "/wo16"], bias=v[a + "/output_transform/bias"]); print("att", rel(att.float(), att_ref.reshape(B * T, D)))
h1 = ops.layernorm_joint_fwd(att, Xr, None, B, T, D, v[a + "/LayerNorm/gamma"], v[a + "/LayerNorm/beta"]); print("h1", rel(h1.float(), h1_ref))
f = ops.gemm(h1.view(B * T, D), sh[a + "/w1_16"], bias=v[a + "/filter_outputencode/bias"], relu=True); print("f", rel(f.float(), f_ref.reshape(B * T, -1)))
bn = a + "/filter_bn"
ops.batch_norm_cols_f16(f, v[bn + "/gamma"], v[bn + "/beta"], v[bn + "/moving_mean"], v[bn + "/moving_variance"], training=False, bessel=False); print("f_bn", rel(f.float(), fbn_ref.reshape(B * T, -1)))
A = ops.gemm(f, sh[a + "/w2_16"], bias=v[a + "/ff_outputencode/bias"], relu=True); print("f2", rel(A.float(), f2_ref.reshape(B * T, K)))
bn = a + "/feed_output_bn"
ops.batch_norm_cols_f16(A, v[bn + "/gamma"], v[bn + "/beta"], v[bn + "/moving_mean"], v[bn + "/moving_variance"], training=False, bessel=False); print("A", rel(A.float(), A_ref.reshape(B * T, K)))
z, rs, a_sum, _ = ops.netvlad_pool_fwd(Xr, B, T, None, None, None, sh["video_VLAD/centers_t"], assign_in=A)
print("a_sum", rel(a_sum, A_ref.sum(1)), "z (unnormalised V^T)", rel(z.float(), Vun.transpose(1, 2)))
# same aggregation from the engine's own A and X in fp64 (isolates the pooling kernel from upstream error)
Vk = torch.matmul(Xr.double().cpu().reshape(B, T, D).transpose(1, 2), A.double().cpu().reshape(B, T, K)) - A.double().cpu().reshape(B, T, K).sum(1, keepdim=True) * C.double()
print("z vs fp64 aggregation of the SAME fp16 inputs", rel(z.float(), Vk.transpose(1, 2)), " |V| typical", float(Vun.abs().mean()), "|A x| typical", float((A_ref.abs().mean()) * xb.abs().mean() * T))
Vn = O.l2_normalize(Vun, 1)                     # [B,D,K] intra
gl = torch.rsqrt(torch.clamp((Vn.reshape(B, -1) ** 2).sum(1), min=1e-12))
rs_ref = torch.rsqrt(torch.clamp((Vun ** 2).sum(1), min=1e-12)) * gl[:, None]      # [B,K]
print("rscale", rel(rs, rs_ref), rs[0, :4].tolist(), rs_ref[0, :4].tolist())
vl = ops.netvlad_finalize(z, rs, d_major=True)
ref_v = O.l2_normalize(Vn.reshape(B, -1), 1)
print("vlad d-major", rel(vl, ref_v))
out16 = torch.zeros(B, D * K, dtype=torch.float16, device=dev)
ops.netvlad_finalize_f16(z, rs, out16, D * K)
print("vlad f16 d-major", rel(out16.float(), ref_v))
# the model-level oracle path
with torch.no_grad():
    full = O.netvlad_atten_cluster_forward(O.batch_norm(xs, P, S, "input_bn", False)[:, :D], P, S, "video_VLAD", T, False)
print("oracle module vs staged oracle", rel(full, ref_v))
