"""Adapters-hub config surface (reference src/configs/adapter_configs.py:1-8).

Upstream maps these names to adapter-transformers config classes, but only ever uses the keys as
``choices=`` for ``--adapter_config`` -- the selected value is never read.  The
names are kept so the CLI parses unchanged without adapter-transformers installed; the values here
describe what the DAT operator actually is (a Pfeiffer-style bottleneck after the FFN output).
"""
ADAPTER_MAP = {
    "pfeiffer": {"placement": "after_ffn_output", "non_linearity": "relu", "reduction_factor": 16},
    "houlsby": {"placement": "after_ffn_output", "non_linearity": "relu", "reduction_factor": 16},
    "parallel": {"placement": "after_ffn_output", "non_linearity": "relu", "reduction_factor": 16},
    "compacter": {"placement": "after_ffn_output", "non_linearity": "relu", "reduction_factor": 16},
}
