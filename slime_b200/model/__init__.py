"""Drop-in module API of the reference's llava/model package, executed by libslime_b200."""
from .language_model.llava_llama import LlavaConfig, LlavaLlamaForCausalLM, LlavaLlamaModel  # noqa: F401
