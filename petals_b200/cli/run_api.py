"""Serve generation over HTTP on top of a running swarm (see petals_b200/client/api_server.py for the API).

    python -m petals.cli.run_api /path/to/model --initial_peers /dev/shm/swarm --device cuda:0 --port 8000
    curl -s localhost:8000/v1/completions -d '{"prompt": [1, 2, 3], "max_tokens": 8}'
"""
from __future__ import annotations

import argparse
import signal

import torch

from petals_b200.client.api_server import ApiServer, GenerationService, load_tokenizer
from petals_b200.constants import DTYPE_MAP
from petals_b200.utils.auto_config import AutoDistributedModelForCausalLM
from petals_b200.utils.logging import get_logger
from petals_b200.utils.paths import resolve_model_path

logger = get_logger(__name__)


def main(argv=None) -> None:
    p = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument("model", help="model name or checkpoint directory (the client reads embeddings / head / tokenizer from it)")
    p.add_argument("--initial_peers", nargs="+", required=True, help="rendezvous directory, inproc://name or /ip4/<host>/tcp/<port> of the swarm")
    p.add_argument("--device", default="cuda:0" if torch.cuda.is_available() else "cpu", help="where the client shell (embeddings, LM head) runs")
    p.add_argument("--torch_dtype", default="auto", choices=sorted(DTYPE_MAP))
    p.add_argument("--host", default="0.0.0.0")
    p.add_argument("--port", type=int, default=8000)
    p.add_argument("--max_session_length", type=int, default=2048, help="tokens of KV cache reserved for a conversation that uses session_id")
    p.add_argument("--session_ttl", type=float, default=300.0, help="seconds after which an idle conversation is closed")
    args = p.parse_args(argv)

    dtype = None if args.torch_dtype == "auto" else DTYPE_MAP[args.torch_dtype]
    model = AutoDistributedModelForCausalLM.from_pretrained(args.model, initial_peers=args.initial_peers, torch_dtype=dtype, device=args.device)
    service = GenerationService(model, load_tokenizer(resolve_model_path(args.model)), model_name=str(args.model),
                                max_session_length=args.max_session_length, session_ttl=args.session_ttl)
    server = ApiServer(service, args.port, args.host)
    logger.info(f"Serving completions on http://{args.host}:{server.port}/v1/completions "
                f"({'text and token-id' if service.tokenizer is not None else 'token-id'} prompts)")
    signal.signal(signal.SIGTERM, lambda *_: (_ for _ in ()).throw(KeyboardInterrupt()))
    try:
        server.serve_forever()
    except KeyboardInterrupt:
        logger.info("shutting down")
    finally:
        server.shutdown()


if __name__ == "__main__":
    main()
