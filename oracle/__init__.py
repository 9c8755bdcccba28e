"""TEST INFRASTRUCTURE.  CPU restatement of the reference's algorithm for the paged-attention decoder
hot path (``golden.py``: torch-native ops; ``store_kv.c``: the integer/byte KV-store work in plain C).
The product package ``mojo_opset_b200`` never imports this."""
