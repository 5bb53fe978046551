"""Prompt assembly in front of the tokenizer (mirrors `groma/data/conversation.py:8-68,70-110` of the
reference: same class name, fields, template names and `get_prompt(messages)` results; pinned by tests/golden/conv_prompts.json).

`messages` is a list of (role, text) pairs -- text may be empty/None for the turn the model has to complete, or a
(text, image, mode) triple whose first element is used -- except for the 'plain' style, which takes bare strings.
"""
from __future__ import annotations

import dataclasses
from typing import Callable, Dict, List, Optional, Sequence


def _text(message):
    return message[0] if isinstance(message, tuple) else message


def _render_single(c: "Conversation", messages) -> str:
    parts = [c.system, c.sep]
    for role, message in messages:
        parts.append(f"{role}: {_text(message)}{c.sep}" if message else f"{role}:")
    return "".join(parts)


def _render_two(c: "Conversation", messages) -> str:
    closers = (c.sep, c.sep2)
    parts = [c.system, c.sep]
    for turn, (role, message) in enumerate(messages):
        parts.append(f"{role}: {_text(message)}{closers[turn % 2]}" if message else f"{role}:")
    return "".join(parts)


def _render_plain(c: "Conversation", messages) -> str:
    closers = (c.sep, c.sep2)
    return c.system + "".join(f"{m}{closers[turn % 2]}" for turn, m in enumerate(messages))


def _render_llama2(c: "Conversation", messages) -> str:
    out = ""
    for turn, (role, message) in enumerate(messages):
        if turn == 0:
            assert message, "first message should not be none"
            assert role == c.roles[0], "first message should come from user"
        if not message:
            continue
        body = _text(message)
        if turn == 0:
            body = f"<<SYS>>\n{c.system}\n<</SYS>>\n\n{body}"
        if turn % 2 == 0:
            out += f"{c.sep}[INST] {body} [/INST]"
        else:
            out += f" {body} {c.sep2}"
    return out.lstrip(c.sep)


_RENDERERS: Dict[str, Callable] = {"single": _render_single, "two": _render_two, "plain": _render_plain, "llama2": _render_llama2}


@dataclasses.dataclass
class Conversation:
    """A prompt template: system text, the two role names and the separators of its style."""
    system: str
    roles: Sequence[str]
    sep_style: str
    sep: str = "###"
    sep2: Optional[str] = None

    def get_prompt(self, messages: List) -> str:
        try:
            render = _RENDERERS[self.sep_style]
        except KeyError:
            raise ValueError(f"Invalid style: {self.sep_style}") from None
        return render(self, messages)


_ASSISTANT_SYSTEM = ("A chat between a curious user and an artificial intelligence assistant. "
                     "The assistant gives helpful, detailed, and polite answers to the user's questions.")

conv_plain = Conversation(system="", roles=("", ""), sep_style="plain", sep=" ", sep2="")
conv_default = Conversation(system=_ASSISTANT_SYSTEM, roles=("USER", "ASSISTANT"), sep_style="two", sep=" ", sep2=" ")
conv_llava = Conversation(system=_ASSISTANT_SYSTEM, roles=("USER", "ASSISTANT"), sep_style="two", sep=" ", sep2="</s>")
conv_llama_2 = Conversation(
    system=("You are a helpful language and vision assistant. You are able to understand the visual content that the user "
            "provides, and assist the user with a variety of tasks using natural language."),
    roles=("USER", "ASSISTANT"), sep_style="llama2", sep="<s>", sep2="</s>")

conv_templates = {"simple": conv_plain, "default": conv_default, "llava": conv_llava, "llama_2": conv_llama_2}
