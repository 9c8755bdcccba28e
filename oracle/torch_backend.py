"""TEST INFRASTRUCTURE - registers the golden restatement as backend ``"torch"``.

Importing this module makes ``MojoXxx._registry.get("torch")`` work exactly as in the reference
(where the core class body *is* the torch backend, ``core/operator.py:31-34``), so the parity tests can be
written in the reference's own idiom::

    op  = MojoPagedDecodeGQA(gqa_layout="AABB")                      # MOJO_BACKEND=b200
    ref = MojoPagedDecodeGQA._registry.get("torch")(gqa_layout="AABB")
    op.forward_diff_with(ref, q, kc, vc, lens, tables, atol=2e-2, rtol=2e-2)

The product never imports it; without it the registry holds only the b200 classes.
"""

from mojo_opset_b200 import core

from . import golden


class TorchPagedDecodeGQA(core.MojoPagedDecodeGQA):
    supported_platforms_list = ["b200", "meta_device"]

    def forward(self, query, key_cache, value_cache, total_seq_lens, block_tables, softmax_scale=None,
                mask=None, *, max_total_seq_len=None):
        core.operators.attention.assert_paged_decode_contract(block_tables, total_seq_lens)
        assert mask is None or self.is_causal, "golden restatement covers the causal path only"
        return golden.paged_decode_gqa(query, key_cache, value_cache, total_seq_lens, block_tables,
                                       softmax_scale, self.gqa_layout)


class TorchPagedPrefillGQA(core.MojoPagedPrefillGQA):
    supported_platforms_list = ["b200", "meta_device"]

    def forward(self, query, key_cache, value_cache, cu_q_lens, block_tables, softmax_scale=None,
                cu_total_seq_lens=None, mask=None, max_q_len=None, max_total_seq_len=None):
        core.operators.attention.assert_paged_prefill_contract(cu_q_lens, block_tables, cu_total_seq_lens)
        assert mask is None, "golden restatement covers mask=None only"
        return golden.paged_prefill_gqa(query, key_cache, value_cache, cu_q_lens, block_tables, softmax_scale,
                                        cu_total_seq_lens, self.gqa_layout, self.is_causal)


class TorchPagedPrefillSWA(core.MojoPagedPrefillSWA):
    supported_platforms_list = ["b200", "meta_device"]

    def forward(self, query, key_cache, value_cache, cu_q_lens, block_table, softmax_scale=None,
                cu_total_seq_lens=None, *, max_q_len=None, max_total_seq_len=None):
        core.operators.attention.assert_paged_prefill_contract(cu_q_lens, block_table, cu_total_seq_lens)
        return golden.paged_prefill_swa(query, key_cache, value_cache, cu_q_lens, block_table, softmax_scale,
                                        cu_total_seq_lens, self.gqa_layout, self.is_causal, self.local_window_size,
                                        self.global_window_size)


class TorchPagedDecodeSWA(core.MojoPagedDecodeSWA):
    supported_platforms_list = ["b200", "meta_device"]

    def forward(self, query, key_cache, value_cache, total_seq_lens, block_table, softmax_scale=None, *,
                max_total_seq_len=None):
        core.operators.attention.assert_paged_decode_contract(block_table, total_seq_lens)
        assert self.is_causal, "golden restatement covers the causal path only"
        return golden.paged_decode_swa(query, key_cache, value_cache, total_seq_lens, block_table, softmax_scale,
                                       self.gqa_layout, self.local_window_size, self.global_window_size)


class TorchSWA(core.MojoSWA):
    supported_platforms_list = ["b200", "meta_device"]

    def forward(self, query, key, value, cu_q_lens, cu_total_seq_lens, softmax_scale=None):
        return golden.swa(query, key, value, cu_q_lens, cu_total_seq_lens, softmax_scale, self.gqa_layout,
                          self.is_causal, self.local_window_size, self.global_window_size)


class TorchSdpa(core.MojoSdpa):
    supported_platforms_list = ["b200", "meta_device"]

    def forward(self, query, key, value, attn_mask=None):
        return golden.sdpa(query, key, value, attn_mask, self.scale, self.enable_gqa)


class TorchStorePagedKVCache(core.MojoStorePagedKVCache):
    supported_platforms_list = ["b200", "meta_device"]

    def forward(self, key_states, value_states, key_cache, value_cache, block_table=None, cu_q_lens=None,
                context_kv_lens=None, *, chunk_metadata=None):
        self._check_store_args(key_states, value_states, block_table, cu_q_lens, context_kv_lens, chunk_metadata)
        if chunk_metadata is None:
            chunk_metadata = golden.build_chunk_plan(block_table, cu_q_lens, context_kv_lens, key_cache.shape[2])
        return golden.store_paged_kv(key_states, value_states, key_cache, value_cache, chunk_metadata)


class TorchRMSNorm(core.MojoRMSNorm):
    supported_platforms_list = ["b200", "meta_device"]

    def forward(self, hidden_state):
        return golden.rms_norm(hidden_state, self.weight, self.variance_epsilon)


class TorchResidualAddRMSNorm(core.MojoResidualAddRMSNorm):
    supported_platforms_list = ["b200", "meta_device"]

    def forward(self, hidden_state, residual):
        return golden.residual_add_rms_norm(hidden_state, residual, self.weight, self.variance_epsilon, self.norm_pos)


class TorchApplyRoPE(core.MojoApplyRoPE):
    supported_platforms_list = ["b200", "meta_device"]

    def forward(self, q, k, cos, sin, head_first=True):
        self._check_rope_args(q, k, cos, sin)
        return golden.apply_rope(q, k, cos, sin, head_first)


class TorchRotaryEmbedding(core.MojoRotaryEmbedding):
    supported_platforms_list = ["b200", "meta_device"]

    def forward(self, x, cu_q_lens=None, total_seq_lens=None, position_ids=None):
        self._check_rotary_args(x, cu_q_lens, total_seq_lens, position_ids)
        pos = golden.rotary_positions(x, cu_q_lens, total_seq_lens, position_ids)
        if self.init_max_length is None:
            return golden.rotary_cos_sin(pos, self.inv_freq, self.attention_scaling)
        return self.cos[pos], self.sin[pos]


class TorchSilu(core.MojoSilu):
    supported_platforms_list = ["b200", "meta_device"]

    def forward(self, x):
        return golden.silu(x)


class TorchSwiGLU(core.MojoSwiGLU):
    supported_platforms_list = ["b200", "meta_device"]

    def forward(self, gate_out, up_out):
        return golden.swiglu(gate_out, up_out, self.swiglu_limit)


class TorchGemmAllReduce(core.MojoGemmAllReduce):
    supported_platforms_list = ["b200", "meta_device"]

    def forward(self, input):
        return golden.gemm_allreduce(input, self.weight, self.bias, self.trans_weight, self.process_group)


class TorchRoPEStoreKV(core.MojoRoPEStoreKV):
    supported_platforms_list = ["b200", "meta_device"]

    def forward(self, q, k, v, cos, sin, key_cache, value_cache, block_table, cu_q_lens, context_kv_lens):
        return golden.norm_rope_store_kv(q, k, v, cos, sin, key_cache, value_cache, block_table, cu_q_lens,
                                         context_kv_lens)[0]


class TorchNormRoPEStoreKV(core.MojoNormRoPEStoreKV):
    supported_platforms_list = ["b200", "meta_device"]

    def forward(self, q, k, v, cos, sin, key_cache, value_cache, block_table, cu_q_lens, context_kv_lens):
        return golden.norm_rope_store_kv(q, k, v, cos, sin, key_cache, value_cache, block_table, cu_q_lens,
                                         context_kv_lens, self.q_weight, self.k_weight, self.variance_epsilon)[0]


class TorchGelu(core.MojoGelu):
    supported_platforms_list = ["b200", "meta_device"]

    def forward(self, x):
        return golden.gelu(x)


class TorchLayerNorm(core.MojoLayerNorm):
    supported_platforms_list = ["b200", "meta_device"]

    def forward(self, hidden_state):
        return golden.layer_norm(hidden_state, self.weight, self.bias, self.variance_epsilon)


class TorchGridRoPE(core.MojoGridRoPE):
    supported_platforms_list = ["b200", "meta_device"]

    def forward(self, x, grid_sizes, freqs_list):
        return golden.grid_rope(x, grid_sizes, freqs_list)
