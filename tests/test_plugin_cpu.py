"""CPU: the plugin artefacts of the drop-in (J4): sfw_plugin.xml, the planner header with the reference's signatures,
the proof binary.  The compile-against-the-unmodified-node check itself is oracle/Makefile's `dropin` target
(__graft_entry__.build() runs it wherever /root/reference exists)."""
import hashlib
import os
import re
import subprocess

import pytest

import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
XML_SHA256 = "b27b40d121a38b4ea393d72e891200ac4ee3b6129f1b2005d78d82cc3b924997"  # of /root/reference/sfw_plugin.xml


def test_plugin_xml_registers_the_reference_class():
    xml = open(os.path.join(ROOT, "sfw_plugin.xml"), "rb").read()
    # pluginlib manifest: library social_force_window_planner, class ...::SFWPlannerNode, base nav2_core::Controller
    assert b'<library path="social_force_window_planner">' in xml
    assert b'type="social_force_window_planner::SFWPlannerNode" base_class_type="nav2_core::Controller"' in xml
    assert hashlib.sha256(xml).hexdigest() == XML_SHA256, "sfw_plugin.xml changed (it must match the reference's)"
    ref = "/root/reference/sfw_plugin.xml"
    if os.path.exists(ref):
        assert xml == open(ref, "rb").read(), "sfw_plugin.xml must stay byte-identical to the reference's"


def test_planner_header_keeps_the_reference_signatures():
    h = open(os.path.join(ROOT, "plugin", "include", "social_force_window_planner", "sfw_planner.hpp")).read()
    flat = re.sub(r"\s+", " ", h)
    for sig in [
        "SFWPlanner(const rclcpp_lifecycle::LifecycleNode::SharedPtr &parent, const std::string name, "
        "std::shared_ptr<SFMSensorInterface> &sensor_iface, const nav2_costmap_2d::Costmap2D &costmap, "
        "std::vector<geometry_msgs::msg::Point> footprint_spec);",
        "bool findBestAction(const geometry_msgs::msg::PoseStamped &global_pose, const geometry_msgs::msg::Twist "
        "&global_vel, geometry_msgs::msg::Twist &cmd_vel);",
        "bool updatePlan(const std::vector<geometry_msgs::msg::PoseStamped> &new_plan);",
        "bool isGoalReached();", "void resetGoal();",
        "visualization_msgs::msg::MarkerArray &getMarkers();",
        "geometry_msgs::msg::Polygon getFootprintPolygon() const",
        "std::vector<geometry_msgs::msg::Point> getFootprint() const",
    ]:
        assert sig in flat, sig
    assert "namespace social_force_window_planner" in h and "#ifndef _SFW_PLANNER_HPP_" in h


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="needs the reference sources")
def test_unmodified_reference_node_compiles_and_links_against_the_plugin():
    """make dropin: reference src/sfw_planner_node.cpp + src/sensor_interface.cpp, zero edits, our planner header."""
    r = subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "dropin"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert ol.have_dropin_node()
    syms = subprocess.run(["nm", "-DC", ol.DROPIN_NODE_SO], capture_output=True, text=True).stdout
    assert "sfw_dropin_node_run" in syms
    assert "social_force_window_planner::SFWPlannerNode::configure" in syms
    assert "social_force_window_planner::SFMSensorInterface::laserCb" in syms
    assert "scoreTrajectory" not in syms, "the reference's planner core must not be in the drop-in"
    undefined = subprocess.run(["nm", "-DCu", ol.DROPIN_NODE_SO], capture_output=True, text=True).stdout
    assert "sfw_score" in undefined and "sfw_create" in undefined  # resolved by libsfw_b200.so at load time
