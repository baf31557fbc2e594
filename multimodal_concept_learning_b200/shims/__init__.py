"""Drop-in replacements for the reference's hot-path functions: same names, same argument
meaning, same return types; the arithmetic runs in libmcl_sm100.so on the GPU.

    reference module                                          shim module
    src/multimodal/token_embedding_analysis.py            ->  shims.token_embedding_analysis
    src/multimodal/token_embedding_analysis_imagenet.py   ->  shims.token_embedding_analysis_imagenet
    random_experiments/multi_token_embedding (notebook)   ->  shims.multi_token
    src/multimodal/mllm.py (LM head + loss)               ->  shims.mllm
    src/multimodal/multimodal_training.py (evaluate)      ->  shims.multimodal_training
    src/vision/vision_training.py (criterion + top-1)     ->  shims.vision_training
"""
