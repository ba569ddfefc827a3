#!/bin/bash
# One recorded attempt to obtain the real faiss (faiss-cpu) on a GPU box ("try once to pin real FAISS").
# The reference neither vendors nor pins faiss (environment.yml has no entry); the build container has no network.
# Outcome (either way) is written to gpurun_out/try_faiss.log; on success the wheel and real-FAISS fixtures come
# back under gpurun_out/faiss_wheel/ and gpurun_out/faiss_golden/.
mkdir -p gpurun_out
LOG=gpurun_out/try_faiss.log
{
  echo "== $(date -u +%FT%TZ) try_faiss on $(hostname)"
  python -c "import faiss; print('already importable', faiss.__version__, faiss.__file__)" 2>&1
  echo "== pip config / index reachability"
  python -m pip config list 2>&1
  for url in https://pypi.org/simple/faiss-cpu/ https://files.pythonhosted.org/ https://conda.anaconda.org/pytorch/linux-64/repodata.json; do
    timeout 20 curl -sS -m 15 -o /dev/null -w "%{http_code} $url\n" "$url" 2>&1 || echo "unreachable $url"
  done
  echo "== local wheel search"
  find / -xdev \( -iname "faiss*" -o -iname "*faiss*.whl" \) -not -path "*/proc/*" -not -path "*/gpurun_out/*" -not -path "*/textreact_b200/*" 2>/dev/null | head -20
  echo "== pip download"
  mkdir -p gpurun_out/faiss_wheel
  timeout 120 python -m pip download --no-deps -d gpurun_out/faiss_wheel faiss-cpu 2>&1 | tail -15
  echo "pip download rc=$?"
  ls -la gpurun_out/faiss_wheel
  if ls gpurun_out/faiss_wheel/*.whl >/dev/null 2>&1; then
    python -m pip install --no-deps --target /tmp/faiss_real gpurun_out/faiss_wheel/*.whl 2>&1 | tail -3
    PYTHONPATH=/tmp/faiss_real python tests/golden/make_faiss_golden.py --out gpurun_out/faiss_golden 2>&1 | tail -20
  else
    echo "RESULT: faiss-cpu could not be obtained on the GPU box (no index reachable, no local wheel)"
  fi
} > "$LOG" 2>&1
tail -30 "$LOG"
