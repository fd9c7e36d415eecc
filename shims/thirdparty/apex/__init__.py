"""Stand-in for the parts of NVIDIA/apex the reference's lamb_amp_opt package imports (apex is not vendored in the reference and is
absent offline): apex.multi_tensor_apply.multi_tensor_applier, used at lamb_amp_opt/fused_lamb/fused_lamb.py:6,167-258."""
