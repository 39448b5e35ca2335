"""Drop-in `lib` package: same module paths as the reference's lib/ for the sampling hot path
(lib.cfg_helper.model_cfg_bank, lib.model_zoo.get_model, lib.model_zoo.ddim.DDIMSampler, ...)."""
