"""The client-facing behavioural contract of the reference, checked in one place on a CPU swarm:
compat properties, RemoteSequential container protocol, session attributes and rollback rules, generate() argument rules,
rejected HF options, deep-prompt broadcasting."""
import pytest
import torch

from petals_b200.client.remote_sequential import RemoteSequential
from petals_b200.utils.auto_config import (AutoDistributedConfig, AutoDistributedModel, AutoDistributedModelForCausalLM,
                                           AutoDistributedModelForSequenceClassification)
from tests.utils import checkpoint, swarm_of


@pytest.fixture(scope="module")
def llama():
    path = checkpoint("llama")
    with swarm_of(path, ["0:4"]) as (swarm, _):
        yield path, swarm


def test_auto_classes_and_compat_properties(llama):
    path, swarm = llama
    lm = AutoDistributedModelForCausalLM.from_pretrained(path, initial_peers=swarm)
    base = AutoDistributedModel.from_pretrained(path, initial_peers=swarm)
    cls = AutoDistributedModelForSequenceClassification.from_pretrained(path, initial_peers=swarm, num_labels=3)
    assert isinstance(lm.model.layers, RemoteSequential) and isinstance(base.layers, RemoteSequential)
    assert lm.transformer is lm.model and lm.model.h is lm.model.layers and lm.model.word_embeddings is lm.model.embed_tokens
    assert lm.model.ln_f == lm.model.norm and callable(lm.model.ln_f)  # the final norm (a bound method of the client shell)
    ids = torch.randint(0, lm.config.vocab_size, (2, 5))
    assert cls(ids).logits.shape == (2, 3) and base(ids).last_hidden_state.shape == (2, 5, lm.config.hidden_size)
    bloom_path = checkpoint("bloom")
    with swarm_of(bloom_path, ["0:4"]) as (bswarm, _):
        bloom = AutoDistributedModelForCausalLM.from_pretrained(bloom_path, initial_peers=bswarm)
        assert isinstance(bloom.transformer.h, RemoteSequential) and bloom.transformer.word_embeddings_layernorm is not None


def test_remote_sequential_container_protocol_and_sessions(llama):
    path, swarm = llama
    config = AutoDistributedConfig.from_pretrained(path, initial_peers=swarm)
    seq = RemoteSequential(config, dht=swarm)
    assert len(seq) == config.num_hidden_layers and len(seq[1:3]) == 2 and len(seq[2]) == 1 and len(list(seq)) == len(seq)
    assert seq.active_session is None
    x = torch.randn(1, 3, config.hidden_size)
    with seq.inference_session(max_length=8) as sess:
        assert seq.active_session is sess and sess.num_blocks == len(seq) and sess.position == 0
        y = seq(x)                       # forward inside a session = a step
        assert y.shape == x.shape and y.dtype == x.dtype and sess.position == 3 and seq.position == 3
        with pytest.raises(AssertionError):
            with seq.inference_session(max_length=8):
                pass                     # no nested sessions
        sess.position = 1                # rollback
        assert sess.position == 1
        with pytest.raises(ValueError):
            sess.position = 5            # cannot move forward by assignment
        with pytest.raises(ValueError, match="Maximum length exceeded"):
            sess.step(torch.randn(1, 8, config.hidden_size))
        with pytest.raises(RuntimeError):
            sess.last_token_id = torch.zeros(1, 1, dtype=torch.long)   # nothing generated yet
    assert seq.active_session is None
    other = seq.inference_session(max_length=4)
    with other, seq.use_session(other):
        assert seq.active_session is other


def test_generate_argument_rules_and_rejected_options(llama):
    path, swarm = llama
    model = AutoDistributedModelForCausalLM.from_pretrained(path, initial_peers=swarm)
    ids = torch.randint(0, model.config.vocab_size, (1, 4))
    with pytest.raises(ValueError):
        model.generate(ids)                                      # neither max_length nor max_new_tokens
    with pytest.raises(ValueError):
        model.generate(ids, max_length=8, max_new_tokens=2)      # both
    out = model.generate(ids, max_new_tokens=3, do_sample=0)    # int do_sample accepted (compat)
    assert out.shape == (1, 7) and torch.equal(out[:, :4], ids)
    with model.inference_session(max_length=16) as sess:         # multi-call continuation on one session
        a = model.generate(ids, max_new_tokens=2, session=sess)
        b = model.generate(None, max_new_tokens=2, session=sess)
        assert b.shape[1] == a.shape[1] + 2 and sess.position >= 7 and sess.output_ids is not None
        assert torch.equal(sess.last_token_id, sess.output_ids[:, -1:])
    for bad in (dict(attention_mask=torch.tensor([[1, 0, 1, 1]])), dict(position_ids=torch.tensor([[0, 2, 3, 4]])),
                dict(output_attentions=True), dict(output_hidden_states=True)):
        with pytest.raises((ValueError, NotImplementedError)):
            model(ids, **bad)
    model(ids, attention_mask=torch.ones_like(ids))              # all-ones mask is fine


def test_deep_prompts_broadcast_over_the_batch(llama):
    path, swarm = llama
    config = AutoDistributedConfig.from_pretrained(path, initial_peers=swarm)
    seq = RemoteSequential(config, dht=swarm)
    x = torch.randn(3, 5, config.hidden_size)
    p1 = torch.randn(len(seq), 1, 2, config.hidden_size) * 0.1
    with torch.no_grad():
        assert torch.allclose(seq(x, prompts=p1), seq(x, prompts=p1.expand(-1, 3, -1, -1).contiguous()), atol=1e-5)
        with seq.inference_session(max_length=8) as sess:
            stepped = sess.step(x, prompts=p1)
        assert torch.allclose(stepped, seq(x, prompts=p1), atol=1e-4)


def test_config_options_that_change_the_math_are_not_dropped():
    """A checkpoint option this engine does not implement must be an error at load time, never a silently different model."""
    import pytest

    from petals_b200.models.llama.config import DistributedLlamaConfig
    from petals_b200.models.mixtral.config import DistributedMixtralConfig

    ok = DistributedLlamaConfig(hidden_size=256, num_attention_heads=4, intermediate_size=512)
    assert ok.block_spec().mlp == "swiglu"
    with pytest.raises(NotImplementedError, match="mlp_bias"):
        DistributedLlamaConfig(hidden_size=256, num_attention_heads=4, intermediate_size=512, mlp_bias=True).block_spec()
    with pytest.raises(NotImplementedError, match="hidden_act"):
        DistributedLlamaConfig(hidden_size=256, num_attention_heads=4, intermediate_size=512, hidden_act="gelu").block_spec()
    with pytest.raises(NotImplementedError, match="hidden_act"):
        DistributedMixtralConfig(hidden_size=256, num_attention_heads=4, intermediate_size=512, hidden_act="relu").block_spec()
    with pytest.raises(NotImplementedError, match="rope_scaling"):
        from petals_b200.models.block_oracle import GenericBlock

        spec = DistributedLlamaConfig(hidden_size=256, num_attention_heads=4, intermediate_size=512, rope_scaling={"rope_type": "dynamic", "factor": 2.0}).block_spec()
        GenericBlock(spec, init_std=0.02).rope_cache("cpu")


def test_server_caps_sessions_at_the_rotary_table():
    from petals_b200.parallel.swarm import Swarm
    from petals_b200.server.server import Server
    from tests.utils import checkpoint

    path = checkpoint("llama")  # max_position_embeddings 512 -> tables cover max(512, 2048) positions
    common = dict(initial_peers=Swarm("rotary-limit"), converted_model_name_or_path=path, block_indices="0:1", torch_dtype="float32", device="cpu",
                  throughput=1.0)
    assert Server(inference_max_length=4096, **common).inference_max_length == 2048  # positions past the table would be rotated wrongly
    assert Server(inference_max_length=1024, **common).inference_max_length == 1024
    assert Server(**common).inference_max_length == 2048  # the GQA default of 8192, capped


def test_every_entry_module_imports_first_in_a_fresh_interpreter():
    """Circular imports only show when a particular module happens to be imported first; scripts and tools do import deep modules
    directly. Each of these must load on its own."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    modules = ["petals", "petals_b200.parallel.swarm", "petals_b200.parallel.transport", "petals_b200.parallel.registry",
               "petals_b200.parallel.tensor_parallel", "petals_b200.server.server", "petals_b200.server.stage_engine", "petals_b200.client",
               "petals_b200.utils.dht", "petals_b200.utils.logging", "petals_b200.ops.native", "petals_b200.cli.run_server", "bench", "__graft_entry__"]
    failures = []
    env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
    procs = [(m, subprocess.Popen([sys.executable, "-c", f"import {m}"], cwd=root, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
             for m in modules]
    for m, p in procs:
        _, err = p.communicate(timeout=300)
        if p.returncode != 0:
            failures.append((m, err.strip().splitlines()[-1] if err.strip() else "?"))
    assert not failures, failures


def test_caller_mistakes_fail_fast_instead_of_being_retried():
    """A request that every server would reject is the caller's mistake: it must raise at once. (If it reached a server, the rejection
    would look like a failing server to the retry loop, which re-routes and retries for as long as ``max_retries`` allows — forever by
    default, as in the reference.)"""
    import time

    import pytest
    import torch

    from petals_b200.client.remote_sequential import RemoteSequential
    from petals_b200.utils.auto_config import AutoDistributedConfig
    from tests.utils import checkpoint, swarm_of

    path = checkpoint("llama")
    with swarm_of(path, ["0:4"]) as (swarm, _):
        config = AutoDistributedConfig.from_pretrained(path, initial_peers=swarm)  # max_retries=None: a retried mistake would hang this test
        seq, H = RemoteSequential(config, dht=swarm), config.hidden_size
        t0 = time.monotonic()
        for bad in (torch.randn(1, 2, H + 1), torch.zeros(1, 2, H, dtype=torch.int64), torch.randn(0, 2, H), torch.randn(1, 0, H)):
            with pytest.raises(ValueError):
                seq(bad)
        with pytest.raises(ValueError, match="deep prompts"):
            seq(torch.randn(1, 2, H), prompts=torch.randn(3, 1, 1, H))
        for bad_length in (0, -3, 2.5, True):
            with pytest.raises(ValueError, match="max_length"):
                seq.inference_session(max_length=bad_length).__enter__()
        with torch.inference_mode(), seq.inference_session(max_length=16) as sess:
            assert sess.step(torch.randn(1, 0, H)).shape == (1, 0, H)  # a zero-token step is legal (reference test_full_model.py)
            sess.step(torch.randn(1, 3, H))
            for bad in (torch.randn(2, 1, H), torch.randn(1, 1, H - 1), torch.randn(1, H), torch.zeros(1, 1, H, dtype=torch.int32)):
                with pytest.raises(ValueError):
                    sess.step(bad)
            for bad_hypo in (torch.tensor([0, 0]), torch.tensor([1]), torch.tensor([-1])):
                with pytest.raises(ValueError, match="hypo_ids"):
                    sess.step(torch.randn(1, 1, H), hypo_ids=bad_hypo)
            with pytest.raises(ValueError, match="position"):
                sess.position = 99
            assert sess.position == 3 and sess.step(torch.randn(1, 1, H)).shape == (1, 1, H)  # the session survived all of it
        assert time.monotonic() - t0 < 20


def test_server_arguments_are_validated_with_readable_errors():
    import pytest

    from petals_b200.parallel.swarm import Swarm
    from petals_b200.server.server import Server
    from tests.utils import checkpoint

    common = dict(initial_peers=Swarm("server-args"), converted_model_name_or_path=checkpoint("llama"), torch_dtype="float32", device="cpu", throughput=1.0)
    for bad, needle in [(dict(block_indices="3:1"), "block_indices"), (dict(block_indices="2:9"), "block_indices"), (dict(block_indices="a:b"), "start:end"),
                        (dict(num_blocks=0), "num_blocks"), (dict(num_blocks=99), "num_blocks"), (dict(block_indices="0:1", torch_dtype="int8"), "torch_dtype"),
                        (dict(block_indices="0:1", quant_type="fp4"), "quant_type"), (dict(block_indices="0:1", attn_cache_tokens=-5), "attn_cache_tokens")]:
        with pytest.raises(ValueError, match=needle):
            Server(**dict(common, **bad))
    with pytest.raises(FileNotFoundError):
        Server(**dict(common, converted_model_name_or_path="/nonexistent/model", block_indices="0:1"))
    with pytest.raises(AssertionError, match="not both"):
        Server(**dict(common, block_indices="0:1", num_blocks=1))


def test_generate_rejects_arguments_that_cannot_work():
    import pytest
    import torch

    from petals_b200.utils.auto_config import AutoDistributedModelForCausalLM
    from tests.utils import checkpoint, swarm_of

    path = checkpoint("llama")
    with swarm_of(path, ["0:4"]) as (swarm, _):
        model = AutoDistributedModelForCausalLM.from_pretrained(path, initial_peers=swarm)
        ids = torch.tensor([[1, 2, 3]])
        bad_calls = [dict(), dict(max_length=5, max_new_tokens=2), dict(max_length=2), dict(max_new_tokens=-1), dict(max_new_tokens=2, num_beams=0),
                     dict(max_new_tokens=2, do_sample=True, temperature=0.0), dict(max_new_tokens=2, do_sample=True, top_p=0.0),
                     dict(max_new_tokens=2, do_sample=True, top_k=-1), dict(max_new_tokens=2, num_beams=2, num_return_sequences=3)]
        for kwargs in bad_calls:
            with pytest.raises(ValueError):
                model.generate(ids, **kwargs)
        for bad_ids in (ids.float(), torch.tensor([[model.config.vocab_size]]), torch.tensor([[-1]]), torch.tensor([1, 2, 3])):
            with pytest.raises(ValueError):  # an id outside the embedding table would be an out-of-bounds read on a GPU
                model.generate(bad_ids, max_new_tokens=2)
        assert model.generate(ids, max_new_tokens=0).shape == (1, 3)  # nothing to add is fine
        assert model.generate(ids, max_new_tokens=2, do_sample=True, top_k=0, top_p=1.0).shape == (1, 5)  # 0 / 1.0 disable the filters


def test_server_checks_what_a_client_sends():
    """The server does not rely on clients being well-behaved for anything that indexes device memory."""
    import pytest
    import torch

    from petals_b200.utils.auto_config import AutoDistributedConfig
    from tests.utils import checkpoint, swarm_of

    path = checkpoint("llama")
    with swarm_of(path, ["0:4"]) as (swarm, servers):
        config = AutoDistributedConfig.from_pretrained(path)
        handler = servers[0].module_container.handler
        uids = [f"{config.dht_prefix}.{i}" for i in range(4)]
        H = config.hidden_size
        stream = handler.rpc_inference(uids, {"max_length": 8})
        try:
            stream.step(torch.randn(2, 2, H))
            for bad in (torch.tensor([0, 2]), torch.tensor([-1, 0]), torch.tensor([0]), torch.tensor([0.0, 1.0])):
                with pytest.raises(ValueError, match="hypo_ids"):
                    stream.step(torch.randn(2, 1, H), hypo_ids=bad)
            with pytest.raises(ValueError, match="batch size"):
                stream.step(torch.randn(3, 1, H))
            with pytest.raises(ValueError, match="prompts"):
                stream.step(torch.randn(2, 1, H), prompts=torch.randn(2, 2, 1, H))
            with pytest.raises(ValueError, match="start_from_position"):
                stream.step(torch.randn(2, 1, H), metadata={"start_from_position": 7})
            with pytest.raises(ValueError, match="Maximum length exceeded"):
                stream.step(torch.randn(2, 7, H))
            assert stream.step(torch.randn(2, 1, H), hypo_ids=torch.tensor([1, 0])).shape == (2, 1, H) and stream.position == 3
        finally:
            stream.close()
        with pytest.raises(ValueError):
            handler.rpc_inference(uids, {"max_length": 10 ** 9})
        with pytest.raises(Exception):
            handler.rpc_inference([uids[0], uids[2]], {"max_length": 8})  # not a contiguous chain
        with pytest.raises(Exception):
            handler.rpc_forward(["other-model.0"], torch.randn(1, 1, H))
        # tensors whose shape the kernels index by the model's hidden size
        for bad in (torch.randn(1, 3, H + 2), torch.randn(3, H), torch.zeros(1, 3, H, dtype=torch.int64), torch.randn(0, 3, H)):
            with pytest.raises(ValueError):
                handler.rpc_forward(uids, bad)
            with pytest.raises(ValueError):
                handler.rpc_backward(uids, bad, bad)
        with pytest.raises(ValueError, match="same shape"):
            handler.rpc_backward(uids, torch.randn(1, 3, H), torch.randn(1, 2, H))
        stream = handler.rpc_inference(uids, {"max_length": 8})
        try:
            for bad in (torch.randn(1, 1, H - 2), torch.zeros(1, 1, H, dtype=torch.int32), torch.randn(1, H)):
                with pytest.raises(ValueError):
                    stream.step(bad)
        finally:
            stream.close()


def test_server_to_server_push_is_used_when_consistent_and_dropped_otherwise():
    """Stage i pushes its output to stage i+1 (reference handler.py:320-350, dead on the reference's client side: SURVEY.md §7.4 Q1).
    The successor uses a pushed tensor only if it is exactly what the client sends for that step."""
    import torch

    from petals_b200.utils.auto_config import AutoDistributedConfig
    from tests.utils import checkpoint, swarm_of

    path = checkpoint("llama")
    with swarm_of(path, ["0:4"]) as (swarm, servers):
        config = AutoDistributedConfig.from_pretrained(path)
        handler = servers[0].module_container.handler
        uids = [f"{config.dht_prefix}.{i}" for i in range(4)]
        H = config.hidden_size
        x = torch.randn(1, 3, H)
        other = torch.randn(1, 3, H)
        ref = handler.rpc_inference(uids, {"max_length": 8, "session_id": "ref"})
        expected_x, expected_other = ref.step(x, metadata={"step_id": "a"}), None
        ref.close()
        ref = handler.rpc_inference(uids, {"max_length": 8, "session_id": "ref2"})
        expected_other = ref.step(other, metadata={"step_id": "a"})
        ref.close()

        stream = handler.rpc_inference(uids, {"max_length": 8, "session_id": "s1"})
        handler.rpc_push(uids, other, metadata={"session_id": "s1", "step_id": "step-1"})  # same shape as what the client sends: used
        out = stream.step(x, metadata={"step_id": "step-1"})
        assert torch.allclose(out, expected_other, atol=1e-5) and not torch.allclose(out, expected_x, atol=1e-3)
        stream.close()

        stream = handler.rpc_inference(uids, {"max_length": 8, "session_id": "s2"})
        handler.rpc_push(uids, torch.randn(1, 7, H), metadata={"session_id": "s2", "step_id": "step-1"})  # a replaying predecessor: dropped
        out = stream.step(x, metadata={"step_id": "step-1"})
        assert out.shape == x.shape and torch.allclose(out, expected_x, atol=1e-5) and stream.position == 3
        handler.rpc_push(uids, x, metadata={"session_id": "s2", "step_id": "step-1"})  # late push for a finished step: ignored
        handler.rpc_push(uids, x, metadata={"session_id": "unknown", "step_id": "zzz"})  # unknown session: ignored
        stream.close()
