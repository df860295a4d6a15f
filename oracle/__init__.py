"""CPU oracle for the FedDAT hot path -- TEST INFRASTRUCTURE ONLY.

A numpy restatement of the reference's arithmetic (HaokunChen245/FedDAT, pure PyTorch), pinned
against golden vectors produced by executing the reference's own code (tests/golden/make_golden.py,
run in the build container where /root/reference is mounted).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package; the product (``feddat_b200``) never does and has no CPU fallback.
"""
from .block_oracle import (attention_backward, attention_forward, gelu, gelu_grad, mlp_fc1_gelu,  # noqa: F401
                           mlp_fc2_dgelu)
from .dat_oracle import (adapter_backward, adapter_forward, adapter_layer_forward_bert,  # noqa: F401
                         pack_branches)
from .fedavg_oracle import get_average_net  # noqa: F401
from .mkd_oracle import (albef_answer_loss, bce_with_logits_times_c, kl_loss, mkd_ce_total,  # noqa: F401
                         mkd_total)
