"""dopamine_b200: Dopamine's replay-and-update hot path on B200 (sm_100a).

Only the path named in SURVEY.md section 8 lives here: HBM-resident replay storage,
GPU sum tree, fused batch gather, fused C51 loss/priority — behind the reference's
own Python API (`replay_memory.*`, `agents.rainbow.rainbow_agent`,
`agents.dqn.dqn_agent`), plus the learner built on it (`agents.rainbow.agent`).
"""
__version__ = '0.1.0'
