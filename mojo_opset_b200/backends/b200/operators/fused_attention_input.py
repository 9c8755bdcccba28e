from mojo_opset_b200 import functional as F
from mojo_opset_b200.core import MojoNormRoPEStoreKV
from mojo_opset_b200.core import MojoRoPEStoreKV


class B200RoPEStoreKV(MojoRoPEStoreKV):
    supported_platforms_list = ["b200"]

    def forward(self, q, k, v, cos, sin, key_cache, value_cache, block_table, cu_q_lens, context_kv_lens):
        return F.norm_rope_store_kv(q, k, v, cos, sin, key_cache, value_cache, block_table, cu_q_lens,
                                    context_kv_lens)


class B200NormRoPEStoreKV(MojoNormRoPEStoreKV):
    supported_platforms_list = ["b200"]

    def forward(self, q, k, v, cos, sin, key_cache, value_cache, block_table, cu_q_lens, context_kv_lens):
        return F.norm_rope_store_kv(q, k, v, cos, sin, key_cache, value_cache, block_table, cu_q_lens,
                                    context_kv_lens, self.q_weight, self.k_weight, self.variance_epsilon)
