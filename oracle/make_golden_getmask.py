"""TEST INFRASTRUCTURE ONLY -- records nn.Model.get_mask outputs of the LIVE, UNMODIFIED
reference (nn.py:44-61) into tests/golden/getmask_v1.json.gz.

Run in the build container (needs /root/reference and torch):  python -m oracle.make_golden_getmask

For positions along random MCTS._step walks (the same kind of walk as features_v1) it stores
the position (board, moves) and the 36 bools ``Model().get_mask(torch.tensor(node.to_vector()))``
-- the logits the reference's policy head sets to -inf -- next to the reference's own
``GameState.action_mask()`` (mcts.py:87-91) of the same node.
"""
from __future__ import annotations

import importlib.util
import os
import random

from .make_golden import _dump
from .refload import REFERENCE_ROOT, load_reference


def main():
    import torch

    ns = load_reference()
    spec = importlib.util.spec_from_file_location("qttt_reference_nn", os.path.join(REFERENCE_ROOT, "nn.py"))
    ref_nn = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_nn)                      # the reference file, untouched
    model = ref_nn.Model()
    M = ns.mcts.MCTS
    recs = []
    rng = random.Random(29)
    for _ in range(80):
        mc = M()
        mc.reset(ns.qtttgym.Board(ns.qtttgym.QEvalClassic()))
        node = mc.root
        while True:
            vec = node.to_vector()
            mask = model.get_mask(torch.tensor(vec))
            recs.append({"board": list(node.board), "moves": [list(m) for m in node.moves],
                         "get_mask": [bool(x) for x in mask.tolist()],
                         "action_mask": [bool(x) for x in node.action_mask()]})
            if node.terminal:
                break
            ns.coin.bits.clear()
            ns.coin.feed(0, 1)
            kids = mc._step(node, int(rng.choice(node.actions)))
            node = kids[rng.randrange(len(kids))]
    # batched call shape too: get_mask on a stacked [B, 18, 10] tensor must equal the per-node masks
    _dump("getmask_v1.json.gz", recs)


if __name__ == "__main__":
    main()
