"""Known-answer vectors shared by the oracle tests (CPU) and the CUDA parity tests (GPU).

REFERENCE_* : exact values asserted by the reference's own unit tests (file:line cited).
SURVEY_KATS : SURVEY.md Appendix C — sha256 of .alc and decoded RGB from an independent
              restatement of the reference source (the reference has no golden bitstreams).
"""

# (kind, w, h, f, quality, wavelet, alc_len, sha256(alc), sha256(decoded rgb))
SURVEY_KATS = [
    (0, 4, 4, 2, 90, 0, 3167, "37e038927d2c6b67c5cbfdbc1d70cfcada85664381043e3b23779f7cb9de7fde", "a281c9f49c06b26326be3d965cac299e1bda26b578603bb21e92a406f78ec9c3"),
    (0, 4, 4, 2, 90, 1, 3180, "09558659d98ffbbaa045c3143e5be7d654f731fecbdeae811a5d83341783e3b8", "f0d2a2ed5e4719b011bebb7d568019f91aa424de504e1ba0a0d39a8f3f75ae35"),
    (0, 8, 8, 2, 100, 2, 3342, "9c0f639a5697d4c833a99afc07c4f10b2925e4fe97b91ff819d2fcfdad9719bb", "02d3a4e47113a9667dcb14ec57752365cfa34d82aaa3a82e593371c8a0c2a133"),
    (0, 3, 5, 1, 90, 0, 3158, "c0ed1d0853666cdb6274426d6cfdc9d7817334360a0e6b33ec3a9137c8349719", "7dcffcea41cee2e48bfaa2aeaca7a295cb8d9150ea41bd75c06ae9ad339e93d1"),
    (0, 64, 64, 4, 50, 0, 10241, "c5984e656e8b0d11031f74a4774b3c157e52f09f36b9ac98955527bdf30953a3", "301ec761466f4f66c42db5a77b0cf4ddd960ecfe7ef22fc058fa5664a81f1113"),
    (0, 32, 16, 8, 80, 1, 7473, "8592c10736ed836288b97ed5c407a51ed8f329dd26db8919b4f8bca7fca8d998", "8e3504352dee615a550a561265d3628b8f90f870927b5c0ce9507ca3ca5c6c26"),
    (0, 32, 16, 8, 75, 2, 5509, "c3f5ce0eb3dacb4cf27714d747237c7c70f7eb40242418f3cfd0f1cde81c6e04", "a253453c03bbe9f4ca5653c3fce5ef439c19b2b55d7ceedbf9b400fd8b6907e1"),
    (1, 64, 32, 8, 90, 0, 8889, "487fd4e8f77f08e42757f073105d12cffeeffbf7252d3ef941c2c0fde577c25e", "63d31d9932caea322b95689f2b687ea872b006926d3d51dea683e15028dcb09e"),
    (1, 64, 32, 8, 80, 1, 14270, "52d0b329174a90279be166e1b9caf48fe3a5711c6a610b5d1d729c7be3ef932e", "c77aea8bff0548d610f89bbb8191486cef92269ee7d2930209d7f970f1052efa"),
    (1, 64, 32, 8, 75, 2, 7087, "e981c1f6ba1f97114bc57cc5514410877f6c947d368578e33336839867aca69f", "f0704bb414cb0bbfc65bd3e09200f56cd9cee3ab7841d903cbd9b4779f20593e"),
]

# SURVEY.md Appendix C, fully expanded smallest case: G0 4x4x2, q=90, CDF 5/3
KAT_4x4x2_HEADER_HEX = "414c43430100040000000400000002000000"
KAT_4x4x2_Y_COEFFS = [58, 73, 0, 7, 116, 131, 0, 7, 0, 0, 0, 0, 12, 12, 0, 0,
                      112, 112, 0, 0, 104, 104, 0, 0, 0, 0, 0, 0, -64, -64, 0, 0]
KAT_4x4x2_Y_SYMBOLS = [11, 15, 0, 0, 27, 29, 0, 0, 0, 0, 0, 0, 1, 1, 0, 0,
                       25, 25, 0, 0, 23, 23, 0, 0, 0, 0, 0, 0, 14, 14, 0, 0]
KAT_4x4x2_STREAMS_HEX = ["0179cb25a44b9ae2a947d68b", "076cce5860e7bc88a738", "6e91b4d97d491a"]

# proptest-regressions/wavelet.txt:7-8 and SURVEY.md Appendix C 1-D vectors
W1D_VECTORS = [
    # (wavelet, input, forward)
    (0, [6, 52, 74, -162, -409, -219, -108, 0], [9, 76, -403, -89, 12, 6, 40, 108]),
    (2, [6, 52, 74, -162, -409, -219, -108, 0], [12, 79, -397, -71, 12, 6, 40, 108]),
    (0, list(range(10, 19)), [10, 12, 14, 16, 0, 0, 0, 0, 0]),
]
W1D_ROUNDTRIP_53 = ([6, 52, 74, -162, -409, -219, -108, 0], [6, 52, 74, -161, -409, -218, -107, 1])

# lossless.rs:110-159,178-185 — vectors whose CDF 5/3 round-trip is exact
LOSSLESS_EXACT_1D = [
    [10, 20, 30, 40, 50, 60, 70, 80],
    [42] * 16,
    [0, 255, 0, 255, 0, 255, 0, 255],
    list(range(64)),
    [-100, -50, 0, 50, 100, 150, -200, 200],
    [42],
    [],
    [1, 2, 3, 4, 5, 6, 7, 8],
]
LOSSLESS_EXACT_2D = [(list(range(64)), 8, 8), ([100] * 256, 16, 16)]

# rans.rs:819-830 / SURVEY A.10 worked examples
FREQ_4BIN = ([100, 200, 300, 400], [409, 819, 1228, 1640], [0, 409, 1228, 2456])
