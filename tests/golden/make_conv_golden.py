"""Generate tests/golden/conv_prompts.json by running the REFERENCE's groma/data/conversation.py (pure Python, importable
in the authoring container; /root/reference does not exist on the GPU box, hence the committed fixture)."""
import importlib.util, json, os, sys

spec = importlib.util.spec_from_file_location("ref_conversation", "/root/reference/groma/data/conversation.py")
ref = importlib.util.module_from_spec(spec)
sys.modules["ref_conversation"] = ref
spec.loader.exec_module(ref)

IMG, REG = "<image>", "<region>"
instruct = f"Here is an image with region crops from it. Image: {IMG}. Regions: {REG}."
answer = "Thank you for the image! How can I assist you with it?"
cases = {
    "run_groma": [("USER", instruct), ("ASSISTANT", answer), ("USER", "Describe the image in detail."), ("ASSISTANT", "")],
    "single_turn_open": [("USER", "What is in <refer_box>?"), ("ASSISTANT", None)],
    "multi_turn_closed": [("USER", "a"), ("ASSISTANT", "b"), ("USER", "c"), ("ASSISTANT", "d")],
    "tuple_message": [("USER", ("look at this", "IMAGE_OBJ", "Crop")), ("ASSISTANT", "")],
}
out = {}
for tname, tmpl in ref.conv_templates.items():
    out[tname] = {"fields": {"system": tmpl.system, "roles": list(tmpl.roles), "sep_style": tmpl.sep_style, "sep": tmpl.sep, "sep2": tmpl.sep2}}
    if tmpl.sep_style == "plain":
        out[tname]["prompts"] = {"plain_pair": tmpl.get_prompt([f"{IMG}\n", "a photo of a cat"]),
                                 "plain_four": tmpl.get_prompt(["q1", "a1", "q2", "a2"])}
    else:
        out[tname]["prompts"] = {k: tmpl.get_prompt(v) for k, v in cases.items()}
# the 'single' style is defined by the class but used by no template: pin it through an ad-hoc instance
single = ref.Conversation(system="SYS", roles=("Human", "Assistant"), sep_style="single", sep="###")
out["_single"] = {"fields": {"system": "SYS", "roles": ["Human", "Assistant"], "sep_style": "single", "sep": "###", "sep2": None},
                  "prompts": {k: single.get_prompt(v) for k, v in cases.items()}}
json.dump({"cases": {k: [[r, list(m) if isinstance(m, tuple) else m] for r, m in v] for k, v in cases.items()}, "templates": out},
          open(os.path.join(os.path.dirname(__file__), "conv_prompts.json"), "w"), indent=1)
print("wrote conv_prompts.json")
