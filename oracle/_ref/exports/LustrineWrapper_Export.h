#pragma once
#define LUSTRINE_WRAPPER_EXPORT
