"""GPU: the reference's own CLI driver, UNCHANGED (tests/golden/ref_driver_pyg-mem-tgn.py.txt is a byte copy
of /root/reference/pyg-mem-tgn.py, see make_golden.vendor_driver), executed with runpy on top of this
package's drop-in modules: `utils`, `dependencyGraph`, `neighbor_loader`, `epoch_utils`, `neg_sampler`,
`model_utils` resolve to tgb-tgn-dgl_b200/ (pyg-mem-tgn.py:16-24), the data set is the offline synthetic
`tgbl-wiki@20000`, the config is the reference's schema with one epoch.  The script's own prints
(pyg-mem-tgn.py:59,63,67) are parsed: finite training loss, validation MRR in (0, 1]."""
import io
import math
import os
import re
import runpy
import shutil
import sys
from contextlib import redirect_stdout

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(os.path.dirname(HERE), "tgb-tgn-dgl_b200")

CONFIG = """sampling:
  - layer: 1
    neighbor:
      - 10
    strategy: 'recent'
    prop_time: False
    history: 1
    duration: 0
    num_thread: 32
memory:
  - type: 'node'
    dim_time: 100
    deliver_to: 'self'
    mail_combine: 'last'
    memory_update: 'gru'
    mailbox_size: 1
    combine_node_feature: True
    dim_out: 100
gnn:
  - arch: 'transformer_attention'
    layer: 1
    att_head: 8
    dim_time: 100
    dim_out: 100
train:
  - epoch: 1
    batch_size: 200
    lr: 0.0001
    dropout: 0.2
    att_dropout: 0.2
    all_on_gpu: True
"""


def test_reference_driver_runs_unchanged(tmp_path):
    script = tmp_path / "pyg-mem-tgn.py"
    shutil.copyfile(os.path.join(HERE, "golden", "ref_driver_pyg-mem-tgn.py.txt"), script)
    cfg = tmp_path / "TGN.yml"
    cfg.write_text(CONFIG)
    argv, path = sys.argv, list(sys.path)
    sys.argv = [str(script), "--data", "tgbl-wiki@20000", "--config", str(cfg)]
    sys.path.insert(0, PKG)
    buf = io.StringIO()
    try:
        torch.manual_seed(0)
        with redirect_stdout(buf):
            runpy.run_path(str(script), run_name="__main__")
    finally:
        sys.argv, sys.path[:] = argv, path
    out = buf.getvalue()
    m_loss = re.search(r"Epoch: 01, Loss: ([-0-9.eE+naninf]+), Training elapsed Time", out)
    m_val = re.search(r"Validation (\w+):\s+([-0-9.eE+naninf]+), elapsed Time", out)
    assert m_loss and m_val and "Execution Time" in out, out[-2000:]
    loss, mrr = float(m_loss.group(1)), float(m_val.group(2))
    n_train = int(20000 * 0.70)
    assert math.isfinite(loss) and 0.0 < loss / n_train < 2.0, out[-2000:]      # sum of loss * batch (epoch_utils.py:310,318)
    assert m_val.group(1) == "mrr" and 0.0 < mrr <= 1.0, out[-2000:]
    # an untrained model on 20 negatives scores ~0.17 (harmonic mean rank); one epoch must do better
    assert mrr > 0.2, out[-2000:]
