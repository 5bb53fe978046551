"""Prompt templates (groma.data.conversation) against prompts produced by the reference module (tests/golden/conv_prompts.json)."""
import json
import os

import pytest

from groma.data.conversation import Conversation, conv_templates

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "conv_prompts.json")))


def _msgs(case):
    return [(r, tuple(m) if isinstance(m, list) else m) for r, m in case]


def test_template_names_and_fields_match_reference():
    assert set(conv_templates) == {k for k in GOLD["templates"] if not k.startswith("_")}
    for name, t in conv_templates.items():
        f = GOLD["templates"][name]["fields"]
        assert (t.system, list(t.roles), t.sep_style, t.sep, t.sep2) == (f["system"], f["roles"], f["sep_style"], f["sep"], f["sep2"])


@pytest.mark.parametrize("name", sorted(GOLD["templates"]))
def test_prompts_match_reference(name):
    g = GOLD["templates"][name]
    f = g["fields"]
    t = conv_templates.get(name) or Conversation(system=f["system"], roles=tuple(f["roles"]), sep_style=f["sep_style"], sep=f["sep"], sep2=f["sep2"])
    for key, want in g["prompts"].items():
        if key == "plain_pair":
            got = t.get_prompt(["<image>\n", "a photo of a cat"])
        elif key == "plain_four":
            got = t.get_prompt(["q1", "a1", "q2", "a2"])
        else:
            got = t.get_prompt(_msgs(GOLD["cases"][key]))
        assert got == want, (name, key)


def test_invalid_style_raises_value_error():
    with pytest.raises(ValueError):
        Conversation(system="", roles=("a", "b"), sep_style="nope").get_prompt([("a", "x")])
