"""Training step (clip + Adam on the flat buffer) vs the oracle's restatement of the reference trainer
(utils.py:170-213, train.py:321-336) on identical weights / batch; gating off isolates the step from the
batch-of-4 gating-BN conditioning (see test_backward_gpu)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_train_step_matches_oracle(cuda):
    from learnablepoolingmethods_b200 import variables
    from learnablepoolingmethods_b200.engine import NetVladConfig, NetVladEngine
    from learnablepoolingmethods_b200.trainer import Trainer
    from oracle import netvlad_oracle as O
    from tests.helpers import oracle_params, perturb, rel
    B, K, Hd, V, T = 4, 64, 64, 100, 128
    store = variables.VariableStore(cuda, seed=11)
    cfg = NetVladConfig(model="NetVladV1", iterations=T, cluster_size=K, hidden_size=Hd, vocab_size=V, gating=False)
    eng = NetVladEngine(cfg, store)
    perturb(store, seed=5)
    P, S = oracle_params(store)
    for p in P.values():
        p.requires_grad_(True)
    tr = Trainer(eng, base_learning_rate=2e-4, batch_size=B)
    opt_state = {}
    model_fn = lambda x, Pp, Ss: O.netvlad_v1(x[0], x[1], Pp, Ss, vocab_size=V, iterations=T, cluster_size=K,
                                              is_training=True, gating=False)
    for step in range(3):
        x, nf, labels = O.synthetic_batch(B, seed=100 + step, vocab=V)
        losses, _ = O.train_step(model_fn, P, S, opt_state, [(x, nf)], [labels], step=step + 1, lr=2e-4)
        loss = tr.train_step(x.to(cuda), nf.to(cuda), labels.to(torch.uint8).to(cuda))
        assert abs(float(loss) - losses[0]) / losses[0] < 1e-2, (step, float(loss), losses[0])
    assert not tr.overflowed()
    # Adam's first steps move every weight by ~lr regardless of gradient scale: compare the UPDATE direction
    worst = 0.0
    P0, _ = oracle_params(variables.VariableStore("cpu", seed=11)) if False else (None, None)
    for name, p in P.items():
        e = rel(store.vars[name], p.detach())
        worst = max(worst, e)
        assert e < 2e-3, (name, e)
    print(f"\n[train 3 steps] worst parameter rel-L2 vs oracle {worst:.2e}")
    assert tr.global_step == 3
