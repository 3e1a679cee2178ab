"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's algorithms for the hot path (ironmanaudi/GNN-decode), used as the checker by tests/,
__graft_entry__.smoke() and bench.py's `cpu_baseline` / `--impl reference` legs.  Nothing under gnn_decode_b200/ imports this
package; the product path has no CPU fallback.

  shim/, ref_loader.py  run the reference's own, unmodified class bodies on CPU (stand-ins for torch_geometric / torch_scatter /
                        matplotlib, which the reference imports but this image lacks) -- the pin for everything below
  restate.py            the seven phase programs (decoder_v2_4, CGNNI, QGNNI, quantum/BP, classical/BP, neural_BP, QGNNNI_ca),
                        forward and the decoder_v2_4 training loss / gradients; bit-identical to the shimmed reference
                        (tests/test_oracle.py) -- parity status: PINNED to the reference run in this container
  philox.py             the syndrome sampler's distribution / layout (error_generate.gen_syn, CGNNI.Gen_Data) on Philox4x32-10
                        and the failure counters (neural_BP.LossFunc, train=0)
  adam.py               torch.optim.Adam's step (the reference's optimizer), pinned against torch.optim.Adam (tests/test_adam.py)
  tables.py             the cubic Hermite tables of the decoder_v2_4 kernels and their error bound (tests/test_tables.py)
  make_golden.py        wrote tests/golden/*.npz from the reference's classes with the shipped checkpoints and seeded inputs
"""
