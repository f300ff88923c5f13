# petals_b200 stage/client image (reference: Dockerfile:1-31, CUDA 11 base). Blackwell needs CUDA >= 12.8.
FROM nvidia/cuda:12.9.0-devel-ubuntu24.04
LABEL repository="petals_b200"
WORKDIR /home
ENV LC_ALL=C.UTF-8 LANG=C.UTF-8 DEBIAN_FRONTEND=noninteractive
RUN apt-get update && apt-get install -y --no-install-recommends python3 python3-pip python3-venv build-essential git && rm -rf /var/lib/apt/lists/*
RUN python3 -m venv /opt/venv && /opt/venv/bin/pip install --no-cache-dir torch numpy msgpack pydantic pyyaml safetensors pytest
ENV PATH="/opt/venv/bin:${PATH}"
VOLUME /cache
ENV PETALS_CACHE=/cache
COPY . petals_b200/
# compile every kernel for sm_100a at image build time (nvcc cross-compiles without a GPU)
RUN cd petals_b200 && python -c "import __graft_entry__ as g; g.build()"
WORKDIR /home/petals_b200/
ENV PYTHONPATH=/home/petals_b200
CMD ["bash"]
